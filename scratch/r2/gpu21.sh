#!/bin/bash
# round 2, call 21: chained job scan; fill<4> source profile with lane-major records; the other workloads on this build
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "cigar or cudamalloc or config0 or multi_chunk" > $OUT/r2_21_pytest.log 2>&1; tail -2 $OUT/r2_21_pytest.log
j() { python - <<PY
import json
d=json.loads(open("gpurun_out/r2_21_$1.json").read().strip().splitlines()[-1])
print("$1:", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["int32_roofline"]["extend"].get("gcups"), d.get("latency_ms"), (d.get("cpu_baseline") or {}).get("value"))
if "mapping_only" in d: print("   MO", round(d["mapping_only"]["value"]), round(d["mapping_only"]["e2e"]["value"]), d["mapping_only"].get("latency_ms"))
PY
}
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > $OUT/r2_21_human.json 2> $OUT/r2_21_human.err; tail -2 $OUT/r2_21_human.err; j human
timeout 900 python bench.py --workload human-repeats --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r2_21_repeats.json 2> $OUT/r2_21_repeats.err; tail -2 $OUT/r2_21_repeats.err; j repeats
timeout 900 python bench.py --workload hifi --ref human --reads 20000 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/r2_21_hifi.json 2> $OUT/r2_21_hifi.err; tail -2 $OUT/r2_21_hifi.err; j hifi
timeout 900 python bench.py --workload prefix --ref human --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r2_21_prefix.json 2> $OUT/r2_21_prefix.err; tail -2 $OUT/r2_21_prefix.err; j prefix
timeout 900 python bench.py --workload config1 --steps 2 --warmup 1 --cpu-sample 4000 > $OUT/r2_21_config1.json 2> $OUT/r2_21_config1.err; tail -2 $OUT/r2_21_config1.err; j config1
MMG_BENCH_PROFILER_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ext_fill_kernel" --launch-count 2 -o $OUT/r2_21_fill -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > $OUT/r2_21_ncu.log 2>&1
ncu -i $OUT/r2_21_fill.ncu-rep --page raw --csv > $OUT/r2_21_fill_raw.csv 2>/dev/null
ncu -i $OUT/r2_21_fill.ncu-rep --page source --csv --print-source sass > $OUT/r2_21_fill_sass.csv 2>/dev/null
rm -f $OUT/r2_21_fill.ncu-rep
