#!/bin/bash
# round 2, call 3: ncu of the register-resident fill kernel and the general DP kernel (config1, CIGAR on, 20k reads)
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --workload config1 --reads 20000 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2_03_bench_c1.json 2> gpurun_out/r2_03_bench_c1.err; tail -2 gpurun_out/r2_03_bench_c1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_03_bench_c1.json").read().strip().splitlines()[-1])
print("c1 cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d["counters"].get("n_cell_fill"), "gap", d["host_gap_ms_per_step"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ext_fill_kernel|ext_dp_kernel" --launch-count 6 -o gpurun_out/r2_03_ext -f python bench.py --workload config1 --reads 20000 --steps 1 --warmup 0 --no-secondary --no-cpu-baseline > gpurun_out/r2_03_ncu.log 2>&1
tail -3 gpurun_out/r2_03_ncu.log | cut -c1-200
ls -la gpurun_out/r2_03_ext.ncu-rep
