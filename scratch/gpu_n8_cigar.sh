set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --cigar --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n8_cigar.json 2> gpurun_out/bench_n8_cigar.err; tail -3 gpurun_out/bench_n8_cigar.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n8_cigar.json").read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["config"]["index_build_s"], d["clocks"])
PY
