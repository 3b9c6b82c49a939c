#!/bin/bash
# round 2, call 1: all GPU tests (new config-size parity tests), smoke, the two-mode default bench line, the CPU arm
cd $GRAFT_REPO_ROOT
nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_01_smoke.log 2>&1; tail -3 gpurun_out/r2_01_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2_01_pytest.log 2>&1; tail -25 gpurun_out/r2_01_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err; tail -3 gpurun_out/r2_01_bench.err; cut -c1-600 gpurun_out/r2_01_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_01_ref.json 2> gpurun_out/r2_01_ref.err; cut -c1-300 gpurun_out/r2_01_ref.json
