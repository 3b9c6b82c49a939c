set -x
cd $GRAFT_REPO_ROOT
for sm in 512 256; do
MMG_SORT_SMALL_MAX=$sm timeout 1200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_h_sm$sm.json 2> gpurun_out/bench_h_sm.err; tail -2 gpurun_out/bench_h_sm.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_h_sm$sm.json").read().strip().splitlines()[-1])
print("small_max $sm", d["value"], d["e2e"]["value"], d["stage_ms_per_step"])
PY
done
