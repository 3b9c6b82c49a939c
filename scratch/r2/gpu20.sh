#!/bin/bash
# round 2, call 20: lane-major traceback records vs row-major; ordered job scan with 8 jobs per thread; Python API
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "cigar or cudamalloc or config0 or multi_chunk" > $OUT/r2_20_pytest.log 2>&1; tail -2 $OUT/r2_20_pytest.log
run() {
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > $OUT/r2_20_$1.json 2> $OUT/r2_20_$1.err; tail -2 $OUT/r2_20_$1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_20_$1.json").read().strip().splitlines()[-1])
print("$1: human cigar", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["int32_roofline"]["extend"]["gcups"])
PY
}
run lane_major
timeout 600 python scratch/api_bench.py > $OUT/r2_20_api.log 2>&1; tail -4 $OUT/r2_20_api.log | cut -c1-200
MMG_BENCH_PROFILER_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/r2_20_launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > $OUT/r2_20_ncu_list.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/r2_20_launches.csv') if not l.startswith('==')]
tot=collections.Counter(); cnt=collections.Counter()
for row in csv.DictReader(lines):
    n=row['Kernel Name'].split('(')[0][:40]; tot[n]+=float(row['Metric Value'].replace(',',''))/1e6; cnt[n]+=1
for n,v in tot.most_common(8): print(f"{n:42s} {cnt[n]:5d} {v:9.1f} ms")
PY
sed -i 's/#define FILL_LANE_MAJOR 1/#define FILL_LANE_MAJOR 0/' mappy-rs_b200/csrc/extend_fill.inc
touch mappy-rs_b200/csrc/extend.cu; make -C mappy-rs_b200 -j16 > $OUT/r2_20_make.log 2>&1
run row_major
