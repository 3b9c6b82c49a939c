#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q -k "config1 or repeats or multi_chunk or isolated" > gpurun_out/pytest_gpu39.log 2>&1; tail -1 gpurun_out/pytest_gpu39.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b39.json 2> gpurun_out/b39.err; tail -2 gpurun_out/b39.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b39.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
