#!/bin/bash
# round 2, call 15: profiles recipe at the current HEAD + the repeat-stress workload
cd $GRAFT_REPO_ROOT
timeout 1500 make -C profiles r02 > gpurun_out/r2_15_make.log 2>&1; tail -3 gpurun_out/r2_15_make.log
timeout 900 python bench.py --workload human-repeats --steps 2 --warmup 1 --cpu-sample 2000 > gpurun_out/r2_15_repeats.json 2> gpurun_out/r2_15_repeats.err; tail -2 gpurun_out/r2_15_repeats.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_15_repeats.json").read().strip().splitlines()[-1])
print("repeats", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"], d["cpu_baseline"])
print("MO", round(d["mapping_only"]["value"]), d["mapping_only"].get("stage_ms_per_step"))
PY
ls -la gpurun_out | head -40; du -sh gpurun_out
