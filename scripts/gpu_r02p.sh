#!/bin/bash
OUT=gpurun_out/${1:-r02p}
mkdir -p $OUT
for n in 5 3 2 1; do
  AB2_SRC_CTAS_PER_SM=$n timeout 300 python bench.py --workload decoder --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 > $OUT/bench_decoder_ctas$n.json 2> $OUT/err.txt
  AB2_SRC_CTAS_PER_SM=$n timeout 300 python bench.py --workload processor --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 > $OUT/bench_processor_ctas$n.json 2>> $OUT/err.txt
  python -c "
import json
for w in ('decoder','processor'):
    d=json.loads(open('$OUT/bench_%s_ctas$n.json' % w).read().strip().splitlines()[-1]); print($n, w, round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['kernels'].items()})"
done
timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 1500 $OUT/bench.json; tail -3 $OUT/bench.err | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_gtconv.py tests/test_gpu_graphconv_blocks.py -q -m gpu > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log | cut -c1-300
