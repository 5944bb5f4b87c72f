#!/bin/bash
# r02u: segmented-A GEMM + multi-linear + LN fork (tests), split-K work-item order (probe), model step with / without recompute
OUT=gpurun_out/${1:-r02u}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_abi.py -q -m gpu -x > $OUT/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -5 $OUT/pytest_gemm.log | cut -c1-300
timeout 600 python scripts/gemm_probe.py --only-perf > $OUT/gemm_probe.jsonl 2>&1
python - <<PY
import json
for l in open('$OUT/gemm_probe.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'tflops' in d: print(d['case'], round(d['ms'],3), round(d['tflops']), round(d['cublas_tflops']), d.get('ok'))
    elif 'error' in d: print(d['case'], 'ERROR', d['error'][-200:])
PY
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; tail -c 300 $OUT/bench_model.err
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --model-recompute off > $OUT/bench_model_norecompute.json 2> $OUT/bench_model_norecompute.err; tail -c 300 $OUT/bench_model_norecompute.err
python - <<PY
import json
for f in ('bench_model','bench_model_norecompute'):
    try:
        d=json.loads(open('$OUT/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],2), d['peak_mem_GB'], d['config']['loss'], d['clocks'])
        for x in (d.get('kernel_breakdown') or [])[:16]: print('   ', round(x['ms'],2), x.get('calls'), x['kernel'][:100])
    except Exception as e: print(f, 'ERR', e)
PY
