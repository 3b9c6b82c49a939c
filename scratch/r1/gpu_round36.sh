#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu36.log 2>&1; tail -2 gpurun_out/pytest_gpu36.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; tail -2 gpurun_out/final2_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final2_bench.json").read().strip().splitlines()[-1])
print("default", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"])
PY
timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cigar36.json 2> gpurun_out/bench_cigar36.err
timeout 900 python bench.py --workload hifi --reads 8000 --cigar --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_hifi_cigar36.json 2> gpurun_out/bench_hifi_cigar36.err
python - <<'PY'
import json
for f in ("bench_cigar36","bench_hifi_cigar36"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/cigar36_launches.csv python bench.py --workload config1 --cigar --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list36.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ext_dp_kernel|ext_stitch_kernel" --launch-skip 2 --launch-count 2 -o gpurun_out/prof_r36_ext -f python bench.py --workload config1 --cigar --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r36.log 2>&1
tail -2 gpurun_out/ncu_r36.log | cut -c1-200
