#!/bin/bash
OUT=gpurun_out/${1:-r02au}
mkdir -p $OUT
for c in perf_kv perf_wgrad perf_mlp2; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $OUT/ncu_gemm_$c python scripts/gemm_probe.py --case $c > $OUT/ncu_$c.log 2>&1
done
ls $OUT
