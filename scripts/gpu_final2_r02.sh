#!/bin/bash
OUT=gpurun_out/${1:-r02final2}
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log | cut -c1-300
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
print('headline', round(d['ms_per_step'],4), round(d['value']/1e6,1), 'e2e', d['e2e'].get('ms_per_step'), 'roof', d['roofline']['frac'], d['roofline_step']['frac'], 'launches', d['gpu_launches'], d['clocks'])
m=d.get('model_step_n320_o96') or {}
print('model', m.get('ms_per_step'), (m.get('without_activation_checkpointing') or {}).get('ms_per_step'), (m.get('reference_blocks_same_gpu') or {}).get('ms_per_step'))
PY
