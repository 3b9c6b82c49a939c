#!/bin/bash
# round 2, call 16: full ncu capture of the mapping kernels only (profiler range), launch list of the repeat-stress workload
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
MMG_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"ext_fill_kernel|ext_dp_kernel|ext_stitch_kernel|ext_prep_kernel|sketch_kernel|seed_kernel|anchor_filter_kernel|expand_kernel|sort_kernel|chain_dp_kernel|backtrack_kernel|regs_kernel" \
  --launch-count 26 -o $OUT/r02_stages -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > $OUT/r02_ncu_full.log 2>&1
python profiles/summarize.py r02 $OUT/r02_launches.csv $OUT/r02_stages.ncu-rep
python profiles/mk_traffic.py $OUT/r02_stages.ncu-rep "bench.py default (configs[2], CIGAR on), first chunk (96 Mbases) of a step" 1
cp profiles/r02.txt profiles/traffic.json $OUT/
ncu -i $OUT/r02_stages.ncu-rep --page raw --csv > $OUT/r02_stages_raw.csv 2>/dev/null
ncu -i $OUT/r02_stages.ncu-rep --page source --csv --print-source sass -k regex:ext_fill_kernel --launch-count 1 > $OUT/r02_fill_sass.csv 2>$OUT/r02_fill_sass.err || true
ncu -i $OUT/r02_stages.ncu-rep --page source --csv -k regex:ext_fill_kernel --launch-count 1 > $OUT/r02_fill_src.csv 2>>$OUT/r02_fill_sass.err || true
ls -la $OUT/r02_stages.ncu-rep; rm -f $OUT/r02_stages.ncu-rep
MMG_BENCH_PROFILER_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file $OUT/r02_repeats_launches.csv \
  python bench.py --workload human-repeats --steps 1 --warmup 0 --no-cpu-baseline --no-secondary > $OUT/r02_repeats_ncu_list.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/r02_repeats_launches.csv') if not l.startswith('==')]
tot=collections.Counter(); cnt=collections.Counter()
for row in csv.DictReader(lines):
    n=row['Kernel Name'].split('(')[0][:40]; tot[n]+=float(row['Metric Value'].replace(',',''))/1e6; cnt[n]+=1
for n,v in tot.most_common(16): print(f"{n:42s} {cnt[n]:5d} {v:9.1f} ms")
PY
du -sh $OUT
