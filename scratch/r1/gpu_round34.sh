#!/bin/bash
cd /root/repo
timeout 1200 python bench.py --cigar --steps 2 --warmup 1 --cpu-sample 400 > gpurun_out/bench_human_cigar34.json 2> gpurun_out/bench_human_cigar34.err; tail -3 gpurun_out/bench_human_cigar34.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human_cigar34.json").read().strip().splitlines()[-1])
print("human cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d.get("cpu_baseline"))
PY
