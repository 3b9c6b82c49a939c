cd $GRAFT_REPO_ROOT
for tool in memcheck initcheck racecheck; do
  for mode in map cigar; do
    echo "=== $tool $mode"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scratch/sanitize.py $mode > gpurun_out/san_${tool}_${mode}.log 2>&1
    grep -c "Invalid\|Uninitialized\|hazard\|Race" gpurun_out/san_${tool}_${mode}.log; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|ok " gpurun_out/san_${tool}_${mode}.log | head -3
  done
done
