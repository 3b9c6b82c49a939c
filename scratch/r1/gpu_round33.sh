#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu33.log 2>&1; tail -2 gpurun_out/pytest_gpu33.log
timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 2 --warmup 1 --cpu-sample 400 > gpurun_out/bench_cigar33.json 2> gpurun_out/bench_cigar33.err; tail -2 gpurun_out/bench_cigar33.err
timeout 900 python bench.py --workload hifi --reads 8000 --cigar --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_hifi_cigar33.json 2> gpurun_out/bench_hifi_cigar33.err; tail -2 gpurun_out/bench_hifi_cigar33.err
python - <<'PY'
import json
for f in ("bench_cigar33","bench_hifi_cigar33"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d.get("cpu_baseline"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/cigar33_launches.csv python bench.py --workload config1 --cigar --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list33.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ext_dp_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/prof_r33_extdp -f python bench.py --workload config1 --cigar --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r33.log 2>&1
tail -2 gpurun_out/ncu_r33.log | cut -c1-200
