set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; tail -3 gpurun_out/pytest_gpu7.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -2 gpurun_out/bench5.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench5.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"])
PY
