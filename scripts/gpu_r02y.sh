#!/bin/bash
OUT=gpurun_out/${1:-r02y}
mkdir -p $OUT
for ew in 8 16; do
for c in perf_proj perf_proj_bias perf_proj_plain perf_proj_res32 perf_q; do
AB2_GEMM_EW=$ew timeout 120 python scripts/gemm_probe.py --case $c 2>&1 | grep PROBE | sed "s/^PROBE /{\"ew\": $ew, \"rec\": /; s/$/}/" >> $OUT/proj_probe.jsonl
done; done
python - <<PY
import json
for l in open('$OUT/proj_probe.jsonl'):
    d=json.loads(l); r=d['rec']; print(d['ew'], r['case'], r['epi'], round(r['ms'],4), round(r['tflops']), round(r['cublas_tflops']))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $OUT/ncu_gemm_perf_proj python scripts/gemm_probe.py --case perf_proj > $OUT/ncu_perf_proj.log 2>&1
