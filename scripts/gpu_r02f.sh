#!/bin/bash
OUT=gpurun_out/${1:-r02f}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity_r02.py tests/test_gpu_graphconv_blocks.py -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
tail -30 $OUT/pytest.log | cut -c1-300
timeout 600 python scripts/gemm_probe.py --only-perf > $OUT/gemm_probe.jsonl 2> $OUT/gemm_probe.err
cat $OUT/gemm_probe.jsonl | cut -c1-400
timeout 600 python bench.py --workload graphconv --steps 10 --warmup 3 > $OUT/bench_graphconv.json 2> $OUT/bench_graphconv.err; tail -c 1500 $OUT/bench_graphconv.json; tail -3 $OUT/bench_graphconv.err
timeout 900 python bench.py --workload model --steps 5 --warmup 2 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; echo "model exit $?"
tail -c 400 $OUT/bench_model.json; tail -3 $OUT/bench_model.err
