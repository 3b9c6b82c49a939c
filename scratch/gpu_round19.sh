set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --workload hifi --reads 8000 --cigar --steps 1 --warmup 1 --cpu-sample 400 > gpurun_out/bench_hifi_cigar.json 2> gpurun_out/bench_hifi_cigar.err; tail -3 gpurun_out/bench_hifi_cigar.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_hifi_cigar.json").read().strip().splitlines()[-1])
print("hifi cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d.get("cpu_baseline"))
PY
timeout 900 python bench.py --workload hifi --reads 40000 --steps 2 --warmup 1 --cpu-sample 4000 > gpurun_out/bench_hifi.json 2> gpurun_out/bench_hifi.err; tail -3 gpurun_out/bench_hifi.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_hifi.json").read().strip().splitlines()[-1])
print("hifi mapping-only", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d.get("cpu_baseline"))
PY
timeout 900 python bench.py --workload prefix --ref human --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_prefix_human.json 2> gpurun_out/bench_prefix_human.err; tail -3 gpurun_out/bench_prefix_human.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_prefix_human.json").read().strip().splitlines()[-1])
print("prefix human", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d.get("latency_ms"))
PY
