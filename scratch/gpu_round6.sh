set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; tail -3 gpurun_out/pytest_gpu8.log
for sm in 1024 256 0; do
MMG_SORT_SMALL_MAX=$sm python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench6_$sm.json 2> gpurun_out/bench6.err; tail -2 gpurun_out/bench6.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench6_$sm.json").read().strip().splitlines()[-1])
print("small_max $sm", d["value"], d["e2e"]["value"], d["stage_ms_per_step"])
PY
done
timeout 1200 python bench.py --workload human --steps 2 --warmup 2 > gpurun_out/bench_human2.json 2> gpurun_out/bench_human2.err; tail -3 gpurun_out/bench_human2.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human2.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["stage_ms_per_step"], d.get("cpu_baseline"))
PY
