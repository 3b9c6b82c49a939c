set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu13.log 2>&1; tail -6 gpurun_out/pytest_gpu13.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; tail -2 gpurun_out/bench13.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench13.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["gpu_launches"])
PY
timeout 1200 python bench.py --workload human --steps 2 --warmup 2 > gpurun_out/bench_human7.json 2> gpurun_out/bench_human7.err; tail -3 gpurun_out/bench_human7.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human7.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["gpu_launches"], d["counters"], d.get("cpu_baseline"))
PY
