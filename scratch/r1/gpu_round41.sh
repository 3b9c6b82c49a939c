#!/bin/bash
# configs[4] proper: map-hifi with CIGAR on the 3.1 Gb reference
cd /root/repo
timeout 1200 python bench.py --workload hifi --ref human --reads 20000 --cigar --steps 2 --warmup 1 --cpu-sample 100 > gpurun_out/bench_hifi_human_cigar.json 2> gpurun_out/bench_hifi_human_cigar.err; tail -3 gpurun_out/bench_hifi_human_cigar.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_hifi_human_cigar.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d["config"], d.get("cpu_baseline"))
PY
