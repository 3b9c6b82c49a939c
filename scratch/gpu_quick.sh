cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("human", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
