#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu43.log 2>&1; tail -1 gpurun_out/pytest_gpu43.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 > gpurun_out/final5_bench.json 2> gpurun_out/final5_bench.err; tail -2 gpurun_out/final5_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final5_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["sample_matches_gpu"])
PY
