#!/bin/bash
# r02af: LayerNorm backward with the next rows prefetched; EW=8 vs 16 on the activation shapes after the leaner epilogue
OUT=gpurun_out/${1:-r02af}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -x -k "layernorm or fork" > $OUT/pytest_ln.log 2>&1; echo "pytest ln exit $?"; tail -2 $OUT/pytest_ln.log | cut -c1-200
for ew in 8 16; do
for c in perf_mlp1 perf_edge512; do
AB2_GEMM_EW=$ew timeout 120 python scripts/gemm_probe.py --case $c 2>&1 | grep PROBE | sed "s/^PROBE /{\"ew\": $ew, \"rec\": /; s/$/}/" >> $OUT/ew_probe.jsonl
done; done
python - <<PY
import json
for l in open('$OUT/ew_probe.jsonl'):
    d=json.loads(l); r=d['rec']; print(d['ew'], r['case'], r['epi'], round(r['ms'],4), round(r['tflops']), round(r['cublas_tflops']))
PY
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; tail -c 300 $OUT/bench_model.err
AB2_GEMM_EW=8 timeout 900 python bench.py --workload model --steps 5 --warmup 3 > $OUT/bench_model_ew8.json 2> $OUT/bench_model_ew8.err
python - <<PY
import json
for f in ('bench_model','bench_model_ew8'):
    try:
        d=json.loads(open('$OUT/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), d.get('peak_mem_GB'), d['clocks'])
        for x in (d.get('kernel_breakdown') or [])[:26]:
            if 'layernorm' in x['kernel'] or 'colsum' in x['kernel'] or 'ALL' in x['kernel']: print('   ', round(x['ms'],3), x.get('calls'), x['kernel'][:110])
    except Exception as e: print(f, 'ERR', e)
PY
