#!/bin/bash
# round 2, call 17: shared-memory AVL nodes in rechain (repeat-stress workload), CIGAR-mode chunk size, Python API throughput
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "rechain or preset or asm or long_join or scoring or four_tuple" > $OUT/r2_17_pytest.log 2>&1; tail -2 $OUT/r2_17_pytest.log
timeout 900 python bench.py --workload human-repeats --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r2_17_repeats.json 2> $OUT/r2_17_repeats.err; tail -2 $OUT/r2_17_repeats.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_17_repeats.json").read().strip().splitlines()[-1])
print("repeats", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
print("MO", round(d["mapping_only"]["value"]), {k: round(v,1) for k,v in d["mapping_only"]["stage_ms_per_step"].items() if v>0.3})
PY
run() { # tag chunk_mbases
B=$(( $2 << 20 ))
MMG_CHUNK_BASES=$B MMG_CHUNK_READS=524288 MMG_ANCHOR_CAP=$(( $2 * 2 / 3 << 20 )) MMG_REGS_CAP=16777216 MMG_CIGAR_CAP=$(( B * 3 )) MMG_JOBS_CAP=$(( B / 48 )) MMG_KEEP_WORDS=16777216 \
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/r2_17_$1.json 2> $OUT/r2_17_$1.err; tail -2 $OUT/r2_17_$1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_17_$1.json").read().strip().splitlines()[-1])
print("$1", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms_per_step"].items() if v>0.3})
PY
}
run c192 192
run c384 384
timeout 600 python scratch/api_bench.py > $OUT/r2_17_api.log 2>&1; tail -4 $OUT/r2_17_api.log
