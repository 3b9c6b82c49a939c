set -x
cd $GRAFT_REPO_ROOT
nproc; free -g | head -2
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; tail -3 gpurun_out/pytest_gpu6.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -2 gpurun_out/bench4.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench4.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["stage_ms_per_step"], d.get("cpu_baseline"), d["config"], d.get("int32_roofline"))
PY
timeout 1200 python bench.py --workload human --steps 2 --warmup 2 > gpurun_out/bench_human.json 2> gpurun_out/bench_human.err; tail -3 gpurun_out/bench_human.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["stage_ms_per_step"], d.get("cpu_baseline"), d["config"], d["counters"])
PY
timeout 600 python bench.py --workload prefix --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_prefix.json 2> gpurun_out/bench_prefix.err; tail -3 gpurun_out/bench_prefix.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_prefix.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["stage_ms_per_step"], d.get("latency_ms"))
PY
