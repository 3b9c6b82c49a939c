#!/bin/bash
# round 2, call 7: fill kernels at 24 / 16 warps per SM
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -x -k "cigar_mode or config2_cigar or hifi" > gpurun_out/r2_07_pytest.log 2>&1; tail -3 gpurun_out/r2_07_pytest.log
for w in config1; do
timeout 600 python bench.py --workload config1 --reads 20000 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2_07_bench_c1.json 2> gpurun_out/r2_07_bench_c1.err; tail -2 gpurun_out/r2_07_bench_c1.err
done
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_07_bench_c1.json").read().strip().splitlines()[-1])
print("c1 cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d["counters"].get("n_cell_fill"), "gap", d["host_gap_ms_per_step"])
PY
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary --no-cpu-baseline > gpurun_out/r2_07_bench.json 2> gpurun_out/r2_07_bench.err; tail -3 gpurun_out/r2_07_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_07_bench.json").read().strip().splitlines()[-1])
print("cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, "gap", d["host_gap_ms_per_step"], d["int32_roofline"].get("extend"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ext_fill_kernel" --launch-count 2 -o gpurun_out/r2_07_fill -f python bench.py --workload config1 --reads 20000 --steps 1 --warmup 0 --no-secondary --no-cpu-baseline > gpurun_out/r2_07_ncu.log 2>&1
tail -2 gpurun_out/r2_07_ncu.log | cut -c1-200
