#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q -k "multi_chunk or config1_sample or full_size" > gpurun_out/pytest_gpu37.log 2>&1; tail -1 gpurun_out/pytest_gpu37.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b37.json 2> gpurun_out/b37.err; tail -2 gpurun_out/b37.err
python bench.py --workload config1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b37_c1.json 2> gpurun_out/b37_c1.err
timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b37_cg.json 2> gpurun_out/b37_cg.err
python - <<'PY'
import json
for f in ("b37","b37_c1","b37_cg"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["config"].get("batches_per_step"), d["gpu_launches"])
PY
