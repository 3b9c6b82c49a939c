set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu17.log 2>&1; tail -4 gpurun_out/pytest_gpu17.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_human10.json 2> gpurun_out/bench_human10.err; tail -3 gpurun_out/bench_human10.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human10.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["gpu_launches"], d["counters"]["n_dropped"])
PY
python bench.py --workload config1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; tail -2 gpurun_out/bench17.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench17.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["gpu_launches"])
PY
