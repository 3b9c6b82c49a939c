#!/bin/bash
# round 2, call 22: traceback byte taken from the field's high byte, z via multiply-add
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "cigar or cudamalloc or config0 or multi_chunk or four_tuple or preset" > $OUT/r2_22_pytest.log 2>&1; tail -2 $OUT/r2_22_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > $OUT/r2_22_human.json 2> $OUT/r2_22_human.err; tail -2 $OUT/r2_22_human.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_22_human.json").read().strip().splitlines()[-1])
print("human:", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["int32_roofline"]["extend"].get("gcups"))
PY
