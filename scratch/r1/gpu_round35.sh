#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q -k "cigar or python_api or poison or cudamalloc" > gpurun_out/pytest_gpu35.log 2>&1; tail -2 gpurun_out/pytest_gpu35.log
timeout 1200 python bench.py --cigar --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_human_cigar35.json 2> gpurun_out/bench_human_cigar35.err; tail -3 gpurun_out/bench_human_cigar35.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human_cigar35.json").read().strip().splitlines()[-1])
print("human cigar", d["value"], d["e2e"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
python scratch/api_bench.py 2>&1 | tail -3
