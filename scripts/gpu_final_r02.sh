#!/bin/bash
# final validation of the round: full GPU suite, smoke, the driver's two bench arms, launch list of the bench command
OUT=gpurun_out/${1:-r02final}
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log | cut -c1-300
timeout 900 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm exit $?"; tail -c 400 $OUT/bench_reference.json
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
print('headline', round(d['ms_per_step'],4), round(d['value']/1e6,1), 'e2e', d['e2e'].get('ms_per_step'), 'roof', d['roofline']['frac'], d['roofline_step']['frac'], 'launches', d['gpu_launches'], d['clocks'])
m=d.get('model_step_n320_o96') or {}
print('model', m.get('ms_per_step'), (m.get('without_activation_checkpointing') or {}).get('ms_per_step'), (m.get('reference_blocks_same_gpu') or {}).get('ms_per_step'))
print('cpu', (d.get('cpu_baseline') or {}).get('ms_per_step'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-model-step --config5 off > $OUT/bench_under_ncu.log 2>&1; echo "ncu exit $?"; wc -l $OUT/launches_bench.csv
for w in decoder processor; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; done
timeout 300 python bench.py --workload edgepath --steps 10 --warmup 3 > $OUT/bench_edgepath_encoder.json 2> $OUT/bench_edgepath.err
python - <<PY
import json
for w in ('decoder','processor','edgepath_encoder'):
    try:
        d=json.loads(open('$OUT/bench_%s.json' % w).read().strip().splitlines()[-1]); print(w, round(d['ms_per_step'],3), round(d.get('value',0)/1e6,1), {k:d[k] for k in d if k.startswith('ms_')})
    except Exception as e: print(w,'ERR',e)
PY
