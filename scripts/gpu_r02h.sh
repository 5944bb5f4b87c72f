#!/bin/bash
OUT=gpurun_out/${1:-r02h}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log | cut -c1-300
timeout 600 python scripts/gemm_probe.py --only-perf > $OUT/gemm_probe_auto.jsonl 2> $OUT/gemm_probe.err
AB2_GEMM_EW=8 timeout 600 python scripts/gemm_probe.py --only-perf > $OUT/gemm_probe_ew8.jsonl 2>> $OUT/gemm_probe.err
AB2_GEMM_EW=16 timeout 600 python scripts/gemm_probe.py --only-perf > $OUT/gemm_probe_ew16.jsonl 2>> $OUT/gemm_probe.err
timeout 600 python bench.py --workload graphconv --steps 10 --warmup 3 --profile > $OUT/bench_graphconv.json 2> $OUT/bench_graphconv.err; tail -c 300 $OUT/bench_graphconv.err
