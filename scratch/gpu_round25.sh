set -x
cd $GRAFT_REPO_ROOT
for rs in 2 3 4; do
MMG_RAMP_SHIFT=$rs timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_h_ramp$rs.json 2> gpurun_out/bench_h_ramp.err; tail -2 gpurun_out/bench_h_ramp.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_h_ramp$rs.json").read().strip().splitlines()[-1])
print("ramp $rs", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1))
PY
done
MMG_RAMP_SHIFT=3 python bench.py --workload config1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench25.json 2> gpurun_out/bench25.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench25.json").read().strip().splitlines()[-1])
print("config1 ramp3", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1))
PY
