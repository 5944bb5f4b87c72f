#!/bin/bash
# boundary rows on the exchange's side stream: NCCL parity tests + weak-scaling line, A/B against AB2_BOUNDARY_STREAM=0
N=${1:-2}
OUT=gpurun_out/${2:-r02aj}_n$N
mkdir -p $OUT
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi.log
tail -4 $OUT/pytest_multi.log | cut -c1-300
fi
for mode in 1 0; do
AB2_BOUNDARY_STREAM=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$mode bench.py --gpus $N --steps 20 --warmup 5 --config5 off --e2e-steps 0 --weak-split work > $OUT/bench_n${N}_bs$mode.json 2> $OUT/bench_n${N}_bs$mode.err; echo "bench mode $mode exit $?"
python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_n${N}_bs$mode.json').read().strip().splitlines()[-1])
    print('boundary_stream=$mode', 'ms', round(d['ms_per_step'],4), 'G edges/s', round(d['value']/1e9,3), 'parity', d['parity']['max_rel_err'], d['parity']['within_tolerance'], 'equal split', d['equal_count_dst_split'].get('ms_per_step'))
except Exception as e: print('ERR', e); print(open('$OUT/bench_n${N}_bs$mode.err').read()[-1500:])
PY
done
