set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu21.log 2>&1; tail -3 gpurun_out/pytest_gpu21.log
timeout 900 python bench.py --workload config1 --cigar --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cigar21.json 2> gpurun_out/bench_cigar21.err; tail -3 gpurun_out/bench_cigar21.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_cigar21.json").read().strip().splitlines()[-1])
print("cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"])
PY
timeout 900 python bench.py --workload prefix --ref human --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_prefix_human2.json 2> gpurun_out/bench_prefix_human2.err; tail -3 gpurun_out/bench_prefix_human2.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_prefix_human2.json").read().strip().splitlines()[-1])
print("prefix human", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d.get("latency_ms"), d["counters"]["n_dropped"])
PY
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_human11.json 2> gpurun_out/bench_human11.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human11.json").read().strip().splitlines()[-1])
print("human", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
python bench.py --workload config1 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench21.json 2> gpurun_out/bench21.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench21.json").read().strip().splitlines()[-1])
print("config1", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
