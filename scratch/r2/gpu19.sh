#!/bin/bash
# round 2, call 19: multi-block CIGAR slice allocation; source-level profile of ext_fill_kernel<4> after the last changes; Python API
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "cigar or cudamalloc or config0 or multi_chunk" > $OUT/r2_19_pytest.log 2>&1; tail -2 $OUT/r2_19_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > $OUT/r2_19_human.json 2> $OUT/r2_19_human.err; tail -2 $OUT/r2_19_human.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_19_human.json").read().strip().splitlines()[-1])
print("human cigar", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["int32_roofline"]["extend"])
PY
timeout 600 python scratch/api_bench.py > $OUT/r2_19_api.log 2>&1; tail -4 $OUT/r2_19_api.log | cut -c1-200
MMG_BENCH_PROFILER_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ext_fill_kernel" --launch-count 2 -o $OUT/r2_19_fill -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > $OUT/r2_19_ncu.log 2>&1
ncu -i $OUT/r2_19_fill.ncu-rep --page raw --csv > $OUT/r2_19_fill_raw.csv 2>/dev/null
ncu -i $OUT/r2_19_fill.ncu-rep --page source --csv --print-source sass > $OUT/r2_19_fill_sass.csv 2>/dev/null
rm -f $OUT/r2_19_fill.ncu-rep
MMG_BENCH_PROFILER_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/r2_19_launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > $OUT/r2_19_ncu_list.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/r2_19_launches.csv') if not l.startswith('==')]
tot=collections.Counter(); cnt=collections.Counter()
for row in csv.DictReader(lines):
    n=row['Kernel Name'].split('(')[0][:40]; tot[n]+=float(row['Metric Value'].replace(',',''))/1e6; cnt[n]+=1
for n,v in tot.most_common(14): print(f"{n:42s} {cnt[n]:5d} {v:9.1f} ms")
PY
