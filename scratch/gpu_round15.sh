set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu15.log 2>&1; tail -4 gpurun_out/pytest_gpu15.log
timeout 1200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_human9.json 2> gpurun_out/bench_human9.err; tail -3 gpurun_out/bench_human9.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human9.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["gpu_launches"], d["counters"]["n_dropped"])
PY
ncu --set full --clock-control none --import-source on -k regex:"anchor_filter_kernel|expand_kernel|chain_dp_kernel|sketch_kernel" --launch-skip 66 --launch-count 4 -o gpurun_out/prof_r15 -f python bench.py --workload human --ref-bases 1000000000 --reads 60000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r15.log 2>&1
tail -3 gpurun_out/ncu_r15.log | cut -c1-200
