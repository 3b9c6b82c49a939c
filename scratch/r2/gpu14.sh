#!/bin/bash
# round 2, call 14: chaining with 8 / 16 / 32 lanes per read
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x -k "config1 or edge_case or repeats or isolated or multi_chunk or hifi_preset or config2_mapping or config3 or cudamalloc or reverse_complement" > gpurun_out/r2_14_pytest.log 2>&1; tail -3 gpurun_out/r2_14_pytest.log
for L in 8 16 32; do
MMG_CHAIN_LANES=$L timeout 900 python bench.py --steps 3 --warmup 2 --mapping-only --no-cpu-baseline > gpurun_out/r2_14_mo_$L.json 2> gpurun_out/r2_14_mo_$L.err; tail -2 gpurun_out/r2_14_mo_$L.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_14_mo_$L.json").read().strip().splitlines()[-1])
print("lanes $L: MO", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["int32_roofline"].get("chain_dp"))
PY
done
