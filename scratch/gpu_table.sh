cd $GRAFT_REPO_ROOT
for il in 2 4; do
MMG_TABLE_INV_LOAD=$il timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_tab$il.json 2> gpurun_out/bench_tab.err; tail -1 gpurun_out/bench_tab.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_tab$il.json").read().strip().splitlines()[-1])
print("inv_load $il", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["config"]["index_build_s"])
PY
done
