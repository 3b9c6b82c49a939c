#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q -k "cigar" > gpurun_out/pytest_gpu38.log 2>&1; tail -1 gpurun_out/pytest_gpu38.log
for i in 1 2; do timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/b38_cg$i.json 2> gpurun_out/b38_cg$i.err; done
python - <<'PY'
import json
for f in ("b38_cg1","b38_cg2"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2))
PY
