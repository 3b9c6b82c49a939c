set -x
cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:"anchor_filter_kernel|expand_kernel|chain_dp_kernel|sketch_kernel" --launch-skip 19 --launch-count 4 -o gpurun_out/prof_r16 -f python bench.py --workload human --ref-bases 1000000000 --reads 60000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r16.log 2>&1
tail -3 gpurun_out/ncu_r16.log | cut -c1-200
