#!/bin/bash
# round 2, call 8 (2 GPUs): full GPU test-suite incl. the multi-device tests, 2-rank bench with the one-process group leg
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_08_pytest.log 2>&1; tail -14 gpurun_out/r2_08_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_08_bench_n2.json 2> gpurun_out/r2_08_bench_n2.err; tail -3 gpurun_out/r2_08_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_08_bench_n2.json") if l.startswith("{")][-1])
print("N=2 cigar", d["value"], d["e2e"]["value"], d["per_rank_ms_per_step"], "gap", d["host_gap_ms_per_step"])
print("group", d.get("group"))
print("MO", d["mapping_only"]["value"], d["mapping_only"]["e2e"]["value"], d["mapping_only"]["per_rank_ms_per_step"])
PY
