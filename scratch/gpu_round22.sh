set -x
cd $GRAFT_REPO_ROOT
for rs in 2 1 3; do
MMG_RAMP_SHIFT=$rs timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_h_ramp$rs.json 2> gpurun_out/bench_h_ramp.err; tail -2 gpurun_out/bench_h_ramp.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_h_ramp$rs.json").read().strip().splitlines()[-1])
print("ramp $rs", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
done
