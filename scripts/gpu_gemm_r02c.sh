#!/bin/bash
# tcgen05 GEMM: probe (2-CTA default + 1-CTA A/B), launch list and ncu full captures (ours and cuBLAS) of two shapes.
OUT=gpurun_out/${1:-r02c}
mkdir -p $OUT
timeout 900 python scripts/gemm_probe.py --perf --cg1 > $OUT/gemm_probe.jsonl 2> $OUT/gemm_probe.err
tail -3 $OUT/gemm_probe.err
for c in perf_sq8k perf_mlp1 perf_kv; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$c.csv python scripts/gemm_probe.py --case $c > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o $OUT/ncu_${c}_ours python scripts/gemm_probe.py --case $c > $OUT/ncu_${c}_ours.log 2>&1
done
timeout 300 ncu --set full --clock-control none -s 30 -c 2 -f -o $OUT/ncu_perf_sq8k_cublas python scripts/gemm_probe.py --case perf_sq8k > $OUT/ncu_sq8k_cublas.log 2>&1
ls -la $OUT
