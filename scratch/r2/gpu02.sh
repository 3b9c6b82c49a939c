#!/bin/bash
# round 2, call 2: register-resident fill kernel: CIGAR parity tests, default bench, fill kernel on/off
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q -x -k "cigar or config0 or config2 or config3 or four_tuple or cs_md or hifi or cudamalloc" > gpurun_out/r2_02_pytest.log 2>&1; tail -5 gpurun_out/r2_02_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-secondary > gpurun_out/r2_02_bench.json 2> gpurun_out/r2_02_bench.err; tail -3 gpurun_out/r2_02_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_02_bench.json").read().strip().splitlines()[-1])
print("cigar", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d["counters"].get("n_cell_fill"), "gap", d["host_gap_ms_per_step"], d["int32_roofline"].get("extend"), d["cpu_baseline"].get("sample_matches_gpu"))
PY
MMG_FILL_KERNEL=0 timeout 900 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2_02_bench_nofill.json 2> gpurun_out/r2_02_bench_nofill.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_02_bench_nofill.json").read().strip().splitlines()[-1])
print("nofill", d["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
timeout 600 python bench.py --workload hifi --ref human --reads 20000 --steps 2 --warmup 1 --no-secondary --cpu-sample 300 > gpurun_out/r2_02_bench_hifi.json 2> gpurun_out/r2_02_bench_hifi.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_02_bench_hifi.json").read().strip().splitlines()[-1])
print("hifi", d["value"], d["e2e"]["value"], {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d["counters"].get("n_cell_fill"), d["cpu_baseline"])
PY
