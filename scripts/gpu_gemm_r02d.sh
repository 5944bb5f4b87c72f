#!/bin/bash
OUT=gpurun_out/${1:-r02d}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q -m gpu > $OUT/pytest_gemm.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gemm.log
tail -25 $OUT/pytest_gemm.log
timeout 600 python scripts/gemm_probe.py --only-perf --cg1 > $OUT/gemm_probe.jsonl 2> $OUT/gemm_probe.err
tail -3 $OUT/gemm_probe.err
for cg in 1 2; do
  AB2_GEMM_CG=$cg timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o $OUT/ncu_sq8k_cg$cg python scripts/gemm_probe.py --case perf_sq8k > $OUT/ncu_sq8k_cg$cg.log 2>&1
done
AB2_GEMM_CG=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o $OUT/ncu_mlp1_cg1 python scripts/gemm_probe.py --case perf_mlp1 > $OUT/ncu_mlp1.log 2>&1
ls -la $OUT
