#!/bin/bash
# r02z: column-owner LayerNorm backward -- parity, A/B on the model step
OUT=gpurun_out/${1:-r02z}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -x > $OUT/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -3 $OUT/pytest_gemm.log | cut -c1-300
AB2_LN_BWD=rows timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model_lnrows.json 2> $OUT/bench_model_lnrows.err
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; tail -c 300 $OUT/bench_model.err
timeout 900 python bench.py --workload model --steps 5 --warmup 3 --model-recompute off > $OUT/bench_model_norecompute.json 2> $OUT/bench_model_norecompute.err
python - <<PY
import json
for f in ('bench_model_lnrows','bench_model','bench_model_norecompute'):
    try:
        d=json.loads(open('$OUT/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), d.get('peak_mem_GB'), d['clocks'], d['config']['loss'])
        for x in (d.get('kernel_breakdown') or [])[:22]:
            if 'layernorm' in x['kernel'] or 'colsum' in x['kernel'] or 'ALL' in x['kernel']: print('   ', round(x['ms'],2), x.get('calls'), x['kernel'][:100])
    except Exception as e: print(f, 'ERR', e)
PY
