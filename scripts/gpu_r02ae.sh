#!/bin/bash
# r02ae: column-owner LayerNorm forward, colsum with row lanes -- parity, A/B on the model step
OUT=gpurun_out/${1:-r02ae}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity_r02.py -q -m gpu -x > $OUT/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -3 $OUT/pytest_gemm.log | cut -c1-300
AB2_LN_FWD=rows timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model_lnfwdrows.json 2> $OUT/bench_model_lnfwdrows.err
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; tail -c 300 $OUT/bench_model.err
timeout 600 python bench.py --workload graphconv --steps 10 --warmup 3 --profile > $OUT/bench_graphconv.json 2> $OUT/bench_graphconv.err
python - <<PY
import json
for f in ('bench_model_lnfwdrows','bench_model','bench_graphconv'):
    try:
        d=json.loads(open('$OUT/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), d.get('peak_mem_GB'), d['clocks'])
        for x in (d.get('kernel_breakdown') or [])[:26]:
            if f=='bench_graphconv' or 'layernorm' in x['kernel'] or 'colsum' in x['kernel'] or 'ALL' in x['kernel']: print('   ', round(x['ms'],3), x.get('calls'), x['kernel'][:110])
    except Exception as e: print(f, 'ERR', e)
PY
