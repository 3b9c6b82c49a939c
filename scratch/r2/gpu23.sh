#!/bin/bash
# round 2, call 23: re-chaining with the outer range-minimum as a warp scan (tie -> tree replay)
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "rechain or preset or chimera or config2_mapping" > $OUT/r2_23_pytest.log 2>&1; tail -2 $OUT/r2_23_pytest.log
for S in 0 1; do
MMG_RMQ_SERIAL=$S timeout 600 python bench.py --workload human-repeats --mapping-only --steps 2 --warmup 1 --cpu-sample 2000 > $OUT/r2_23_rmq_$S.json 2> $OUT/r2_23_rmq_$S.err; tail -2 $OUT/r2_23_rmq_$S.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_23_rmq_$S.json").read().strip().splitlines()[-1])
print("serial=$S: MO", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_rechain"], (d.get("cpu_baseline") or {}).get("sample_matches_gpu"))
PY
done
