#!/bin/bash
TAG=${1:-ncutma}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
# force the pipelined kernels for every pass on the headline graph and capture them
AB2_TMA=15 AB2_SRC_TMA_ALWAYS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gtconv -s 3 -c 3 -f -o $OUT/gtconv_tma_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ls -la $OUT; tail -2 $OUT/ncu_full.log | cut -c1-300
