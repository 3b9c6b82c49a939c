set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"seed_kernel|anchor_filter_kernel|expand_kernel|sort_kernel|chain_dp_kernel|backtrack_kernel|regs_kernel" --launch-count 9 -o gpurun_out/final_prof_stages -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_ncu_full.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"sketch_kernel" --launch-skip 47 --launch-count 1 -o gpurun_out/final_prof_sketch -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/final_ncu_sketch.log 2>&1
tail -2 gpurun_out/final_ncu_sketch.log | cut -c1-200
