#!/bin/bash
# packed 16x2 ext_dp core: parity (cigar tests) + config1 CIGAR bench + ncu of ext_dp
cd /root/repo
python -m pytest tests -m gpu -x -q -k "cigar or python_api or smoke" > gpurun_out/pytest_gpu32.log 2>&1; tail -3 gpurun_out/pytest_gpu32.log
timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cigar32.json 2> gpurun_out/bench_cigar32.err; tail -3 gpurun_out/bench_cigar32.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_cigar32.json").read().strip().splitlines()[-1])
print("cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"])
PY
ncu --set full --clock-control none --import-source on -k regex:"ext_dp_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/prof_r32_extdp -f python bench.py --workload config1 --cigar --reads 8000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r32.log 2>&1
tail -2 gpurun_out/ncu_r32.log | cut -c1-200
