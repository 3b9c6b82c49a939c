#!/bin/bash
cd /root/repo
timeout 140 python bench.py --workload prefix --ref human --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prefix44.json 2> gpurun_out/bench_prefix44.err; tail -2 gpurun_out/bench_prefix44.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_prefix44.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d.get("latency_ms"), d["config"].get("batches_per_step"))
PY
