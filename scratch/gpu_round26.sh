cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu26.log 2>&1; tail -2 gpurun_out/pytest_gpu26.log
for ds in 1 0; do
MMG_DUAL_STREAM=$ds timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_dual$ds.json 2> gpurun_out/bench_dual.err; tail -2 gpurun_out/bench_dual.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_dual$ds.json").read().strip().splitlines()[-1])
print("human dual=$ds", round(d["value"]), round(d["ms_per_step"],1), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
done
MMG_DUAL_STREAM=1 python bench.py --workload config1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench26.json 2> gpurun_out/bench26.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench26.json").read().strip().splitlines()[-1])
print("config1 dual=1", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
