#!/bin/bash
# round 2, call 18: fill kernel without H tracking + path prefetch + lane-0 table; 384-Mbase chunks with CIGAR; rmq node placement
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "cigar or config0 or four_tuple or cs_md or preset" > $OUT/r2_18_pytest.log 2>&1; tail -2 $OUT/r2_18_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $OUT/r2_18_human.json 2> $OUT/r2_18_human.err; tail -2 $OUT/r2_18_human.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_18_human.json").read().strip().splitlines()[-1])
print("human cigar", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3}, d["int32_roofline"]["extend"], "MO", round(d["mapping_only"]["value"]), round(d["mapping_only"]["e2e"]["value"]))
PY
for cfg in "96 8" "1 8" "256 4" "32 12"; do set -- $cfg
MMG_RMQ_NODES=$1 MMG_RMQ_CTAS=$2 timeout 600 python bench.py --workload human-repeats --mapping-only --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r2_18_rmq_$1_$2.json 2> $OUT/r2_18_rmq_$1_$2.err; tail -2 $OUT/r2_18_rmq_$1_$2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_18_rmq_$1_$2.json").read().strip().splitlines()[-1])
print("rmq nodes $1 ctas $2: MO", round(d["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
done
timeout 900 python bench.py --workload hifi --ref human --reads 20000 --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/r2_18_hifi.json 2> $OUT/r2_18_hifi.err; tail -2 $OUT/r2_18_hifi.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_18_hifi.json").read().strip().splitlines()[-1])
print("hifi", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
