cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_arm.json 2> gpurun_out/ref_arm.err; tail -2 gpurun_out/ref_arm.err; cut -c1-600 gpurun_out/ref_arm.json
