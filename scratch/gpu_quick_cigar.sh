#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q -k "cigar" > gpurun_out/pytest_gpuq.log 2>&1; tail -1 gpurun_out/pytest_gpuq.log
timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cigarq.json 2> gpurun_out/bench_cigarq.err; tail -3 gpurun_out/bench_cigarq.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_cigarq.json").read().strip().splitlines()[-1])
print("cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"])
PY
