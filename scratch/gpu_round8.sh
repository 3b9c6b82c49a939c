set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1; tail -3 gpurun_out/pytest_gpu9.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -2 gpurun_out/bench8.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench8.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"])
PY
timeout 1200 python bench.py --workload human --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_human3.json 2> gpurun_out/bench_human3.err; tail -3 gpurun_out/bench_human3.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human3.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"])
PY
ncu --set full --clock-control none --import-source on -k regex:"sort_kernel" --launch-skip 9 --launch-count 3 -o gpurun_out/prof_r8_sort -f python bench.py --workload human --ref-bases 1000000000 --reads 40000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r8.log 2>&1
tail -3 gpurun_out/ncu_r8.log
