set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --workload human > gpurun_out/bench_n2_human.json 2> gpurun_out/bench_n2_human.err; tail -3 gpurun_out/bench_n2_human.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2_human.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["index_build_s"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; tail -2 gpurun_out/bench_n2_ref.err; cat gpurun_out/bench_n2_ref.json | cut -c1-400
