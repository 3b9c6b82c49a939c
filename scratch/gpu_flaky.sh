cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6; do
  MMG_POISON=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cigar_mode or repeats or config1_stages" > gpurun_out/flaky_$i.log 2>&1
  tail -1 gpurun_out/flaky_$i.log
  grep "^E " gpurun_out/flaky_$i.log | head -6 | cut -c1-600
done
