set -x
cd $GRAFT_REPO_ROOT
python bench.py --workload human --ref-bases 1000000000 --reads 40000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_h1g.json 2> gpurun_out/bench_h1g.err; tail -2 gpurun_out/bench_h1g.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_h1g.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["counters"])
PY
ncu --set full --clock-control none --import-source on -k regex:"sort_kernel|seed_kernel|regs_kernel|expand_kernel|chain_dp_kernel|backtrack_kernel" --launch-skip 7 --launch-count 7 -o gpurun_out/prof_r7_human -f python bench.py --workload human --ref-bases 1000000000 --reads 40000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r7.log 2>&1
tail -3 gpurun_out/ncu_r7.log
