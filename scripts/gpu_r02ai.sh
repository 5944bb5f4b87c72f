#!/bin/bash
# r02ai: L2 tensor prefetch of the residual / saved pre-activation tile by the producer warp -- parity, probe, model
OUT=gpurun_out/${1:-r02ai}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity_r02.py -q -m gpu -x > $OUT/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -3 $OUT/pytest_gemm.log | cut -c1-300
timeout 600 python scripts/gemm_probe.py --only-perf > $OUT/gemm_probe.jsonl 2>&1
python - <<PY
import json
for l in open('$OUT/gemm_probe.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'tflops' in d: print(d['case'], d['epi'], round(d['ms'],4), round(d['tflops']), round(d['cublas_tflops']), d.get('ok'))
    elif 'error' in d: print(d['case'], 'ERROR', d['error'][-200:])
PY
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; tail -c 300 $OUT/bench_model.err
python - <<PY
import json
for f in ('bench_model',):
    try:
        d=json.loads(open('$OUT/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), d.get('peak_mem_GB'), d['clocks'])
        for x in (d.get('kernel_breakdown') or [])[:8]: print('   ', round(x['ms'],2), x.get('calls'), x['kernel'][:100])
    except Exception as e: print(f, 'ERR', e)
PY
