#!/bin/bash
# r02ac: pad/cast rows op; full GPU suite; default bench (driver's command line); model with / without recompute
OUT=gpurun_out/${1:-r02ac}
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 300 $OUT/bench.err
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; tail -c 300 $OUT/bench_model.err
python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
print('headline', round(d['ms_per_step'],4), round(d['value']/1e6,1), 'e2e', d['e2e'].get('ms_per_step'), 'roof', d['roofline']['frac'], d['roofline_step']['frac'])
print('model', d.get('model_step_n320_o96'))
print('cpu', d.get('cpu_baseline'))
d=json.loads(open('$OUT/bench_model.json').read().strip().splitlines()[-1]); print('model', round(d['ms_per_step'],3), d.get('peak_mem_GB'), d['clocks'])
for x in (d.get('kernel_breakdown') or [])[:24]: print('   ', round(x['ms'],2), x.get('calls'), x['kernel'][:100])
PY
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log
