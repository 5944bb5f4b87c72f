#!/bin/bash
# r02w: warp-uniform producer / MMA-issuer loops -- parity tests, probe, ncu of the kv shape, model step
OUT=gpurun_out/${1:-r02w}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -x > $OUT/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -3 $OUT/pytest_gemm.log | cut -c1-300
timeout 600 python scripts/gemm_probe.py --perf > $OUT/gemm_probe.jsonl 2>&1
python - <<PY
import json
for l in open('$OUT/gemm_probe.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    if 'tflops' in d: print(d['case'], round(d['ms'],3), round(d['tflops']), round(d['cublas_tflops']), d.get('ok'))
    elif 'error' in d: print(d['case'], 'ERROR', d['error'][-200:])
    elif not d.get('ok', True): print(d['case'], 'NOT OK', d)
PY
for c in perf_kv perf_sq8k perf_mlp1; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $OUT/ncu_gemm_$c python scripts/gemm_probe.py --case $c > $OUT/ncu_$c.log 2>&1
done
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; tail -c 300 $OUT/bench_model.err
python - <<PY
import json
for f in ('bench_model',):
    try:
        d=json.loads(open('$OUT/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), d.get('peak_mem_GB'), d['clocks'])
        for x in (d.get('kernel_breakdown') or [])[:14]: print('   ', round(x['ms'],2), x.get('calls'), x['kernel'][:100])
    except Exception as e: print(f, 'ERR', e)
PY
