#!/bin/bash
# round 2, call 4: job-pipelined fill kernel: CIGAR parity tests, config1 + default bench, ncu of the fill kernel
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x -k "cigar or config0 or config2 or config3 or four_tuple or cs_md or hifi or cudamalloc" > gpurun_out/r2_06_pytest.log 2>&1; tail -5 gpurun_out/r2_06_pytest.log
timeout 600 python bench.py --workload config1 --reads 20000 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2_06_bench_c1.json 2> gpurun_out/r2_06_bench_c1.err; tail -2 gpurun_out/r2_06_bench_c1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_06_bench_c1.json").read().strip().splitlines()[-1])
print("c1 cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d["counters"].get("n_cell_fill"), "gap", d["host_gap_ms_per_step"])
PY
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary > gpurun_out/r2_06_bench.json 2> gpurun_out/r2_06_bench.err; tail -3 gpurun_out/r2_06_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_06_bench.json").read().strip().splitlines()[-1])
print("cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d["counters"].get("n_cell_fill"), "gap", d["host_gap_ms_per_step"], d["int32_roofline"].get("extend"), d["cpu_baseline"].get("sample_matches_gpu"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ext_fill_kernel" --launch-count 2 -o gpurun_out/r2_06_fill -f python bench.py --workload config1 --reads 20000 --steps 1 --warmup 0 --no-secondary --no-cpu-baseline > gpurun_out/r2_06_ncu.log 2>&1
tail -2 gpurun_out/r2_06_ncu.log | cut -c1-200
