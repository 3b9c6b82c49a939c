set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; tail -3 gpurun_out/pytest_gpu5.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench3.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["stage_ms_per_step"], d.get("cpu_baseline"))
PY
