set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; tail -3 gpurun_out/pytest_gpu10.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -2 gpurun_out/bench9.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench9.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"])
PY
timeout 1200 python bench.py --workload human --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_human4.json 2> gpurun_out/bench_human4.err; tail -3 gpurun_out/bench_human4.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_human4.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"])
PY
timeout 900 python bench.py --cigar --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cigar9.json 2> gpurun_out/bench_cigar9.err; tail -3 gpurun_out/bench_cigar9.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_cigar9.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["counters"])
PY
