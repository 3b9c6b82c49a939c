#!/bin/bash
# round 2, call 24: the profiles recipe at HEAD + smoke + Python API
cd $GRAFT_REPO_ROOT
OUT=gpurun_out
timeout 2400 make -C profiles r02 > $OUT/r2_24_make.log 2>&1; tail -3 $OUT/r2_24_make.log
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value']), round(d['e2e']['value']), d['roofline']['frac'], d['int32_roofline']['extend'], d['cpu_baseline']['value'], d['cpu_baseline'].get('sample_matches_gpu'), 'MO', round(d['mapping_only']['value']), round(d['mapping_only']['e2e']['value']))
r=json.loads(open('gpurun_out/r02_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_24_smoke.log 2>&1; tail -3 $OUT/r2_24_smoke.log
timeout 600 python scratch/api_bench.py > $OUT/r2_24_api.log 2>&1; tail -3 $OUT/r2_24_api.log | cut -c1-160
ls -la $OUT | head -30; du -sh $OUT
