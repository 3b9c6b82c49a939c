set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu23.log 2>&1; tail -3 gpurun_out/pytest_gpu23.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_human12.json 2> gpurun_out/bench_human12.err; tail -2 gpurun_out/bench_human12.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human12.json").read().strip().splitlines()[-1])
print("human", round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
python bench.py --workload config1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench23.json 2> gpurun_out/bench23.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench23.json").read().strip().splitlines()[-1])
print("config1", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cigar23.json 2> gpurun_out/bench_cigar23.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_cigar23.json").read().strip().splitlines()[-1])
print("cigar", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
