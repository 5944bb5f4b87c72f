#!/bin/bash
# multi-GPU validation: NCCL parity tests, then bench.py --gpus N (sharded parity check + config-5 block inside)
N=${1:-2}
OUT=gpurun_out/${2:-r02i}_n$N
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin_model.py -q -m gpu -x > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi.log
tail -25 $OUT/pytest_multi.log | cut -c1-300
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit $?"
tail -c 3500 $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err | cut -c1-400
if [ "${NCCL_BARRIER:-0}" == "1" ]; then
AB2_BARRIER=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 --config5 off --no-parity-check > $OUT/bench_n${N}_ncclbarrier.json 2> $OUT/bench_n${N}_ncclbarrier.err
tail -c 600 $OUT/bench_n${N}_ncclbarrier.json
fi
if [ "${TRACE:-0}" == "1" ]; then
AB2_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 10 --warmup 5 --config5 off --no-parity-check --e2e-steps 0 > $OUT/bench_n${N}_trace.json 2> $OUT/bench_n${N}_trace.err
grep "trace rank" $OUT/bench_n${N}_trace.err | cut -c1-400
fi
