#!/bin/bash
# round 2, call 13: fill kernel: z-drop penalty filter; 4 vs 5 CTAs per SM for ext_fill_kernel<4>; other workloads
cd $GRAFT_REPO_ROOT
run() { # tag
timeout 600 python bench.py --workload config1 --reads 40000 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r2_13_$1.json 2> gpurun_out/r2_13_$1.err; tail -2 gpurun_out/r2_13_$1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_13_$1.json").read().strip().splitlines()[-1])
print("$1 c1 cigar", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
}
run ctas4
sed -i 's/#define FILL_CTAS_4 4 /#define FILL_CTAS_4 5 /' mappy-rs_b200/csrc/extend_fill.inc
touch mappy-rs_b200/csrc/extend.cu; make -C mappy-rs_b200 -j16 > gpurun_out/r2_13_make.log 2>&1; grep -A2 "ext_fill_kernelILi4" mappy-rs_b200/build/extend.ptxas.log | grep registers
run ctas5
timeout 600 python -m pytest tests -m gpu -q -x -k "cigar_mode or config2_cigar" > gpurun_out/r2_13_pytest.log 2>&1; tail -2 gpurun_out/r2_13_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_13_human.json 2> gpurun_out/r2_13_human.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_13_human.json").read().strip().splitlines()[-1])
print("human cigar (ctas5)", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, "MO", round(d["mapping_only"]["value"]), round(d["mapping_only"]["e2e"]["value"]))
PY
timeout 900 python bench.py --workload hifi --ref human --reads 20000 --steps 2 --warmup 1 --no-secondary --cpu-sample 300 > gpurun_out/r2_13_hifi.json 2> gpurun_out/r2_13_hifi.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_13_hifi.json").read().strip().splitlines()[-1])
print("hifi", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["counters"]["n_cell"], d["counters"].get("n_cell_fill"), d["cpu_baseline"]["value"], d["cpu_baseline"]["sample_matches_gpu"])
PY
timeout 900 python bench.py --workload prefix --ref human --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_13_prefix.json 2> gpurun_out/r2_13_prefix.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_13_prefix.json").read().strip().splitlines()[-1])
print("prefix cigar", round(d["value"]), round(d["e2e"]["value"]), d.get("latency_ms"), "MO", round(d["mapping_only"]["value"]), round(d["mapping_only"]["e2e"]["value"]))
PY
