#!/bin/bash
cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu40.log 2>&1; tail -2 gpurun_out/pytest_gpu40.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err; tail -2 gpurun_out/final3_bench.err
python bench.py --workload config1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final3_c1.json 2> gpurun_out/final3_c1.err
timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final3_cg.json 2> gpurun_out/final3_cg.err
python - <<'PY'
import json
for f in ("final3_bench","final3_c1","final3_cg"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"))
PY
