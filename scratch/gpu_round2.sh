set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; tail -3 gpurun_out/pytest_gpu4.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; cat gpurun_out/bench2.json | head -c 1500
python scratch/e2e_diag.py > gpurun_out/e2e_diag2.log 2>&1; tail -12 gpurun_out/e2e_diag2.log
ncu --set full --clock-control none --import-source on -k regex:"sort_kernel|backtrack_kernel|seed_kernel|sketch_kernel|chain_dp_kernel|regs_kernel|expand_kernel" --launch-skip 8 --launch-count 8 -o gpurun_out/prof_r2_all -f python bench.py --steps 1 --warmup 1 --reads 40000 --no-cpu-baseline > gpurun_out/ncu_r2.log 2>&1
tail -3 gpurun_out/ncu_r2.log
