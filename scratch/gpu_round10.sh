set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_human.csv python bench.py --workload human --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r10.log 2>&1
tail -2 gpurun_out/ncu_r10.log | cut -c1-300
