#!/bin/bash
# ncu evidence of round 2 (one GPU): launch list of the default bench command, full captures of the tcgen05 GEMM on two block
# shapes and of the conv kernels on the decoder / processor graphs.
OUT=gpurun_out/${1:-r02n}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --config5 off --no-model-step > $OUT/ncu_launch.log 2>&1
grep -c gtconv $OUT/launches_bench.csv
for c in perf_kv perf_mlp1 perf_edge512; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o $OUT/ncu_gemm_$c python scripts/gemm_probe.py --case $c > $OUT/ncu_gemm_$c.log 2>&1
done
for w in decoder processor; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gtconv -s 3 -c 3 -f -o $OUT/ncu_conv_$w \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > $OUT/ncu_conv_$w.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gtconv -s 3 -c 3 -f -o $OUT/ncu_conv_encoder \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --config5 off --no-model-step > $OUT/ncu_conv_encoder.log 2>&1
ls -la $OUT
