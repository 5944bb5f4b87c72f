#!/bin/bash
OUT=gpurun_out/${1:-r02t}
mkdir -p $OUT
python scripts/pcie_probe.py > $OUT/pcie_probe.json 2>&1; cat $OUT/pcie_probe.json
for il in 0 1; do
  AB2_SRC_INTERLEAVE=$il timeout 300 python bench.py --workload processor --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 > $OUT/bench_processor_il$il.json 2> $OUT/err.txt
  AB2_SRC_INTERLEAVE=$il timeout 300 python bench.py --workload decoder --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 > $OUT/bench_decoder_il$il.json 2>> $OUT/err.txt
  python -c "
import json
for w in ('decoder','processor'):
    d=json.loads(open('$OUT/bench_%s_il$il.json' % w).read().strip().splitlines()[-1]); print($il, w, round(d['ms_per_step'],3), round(d['value']/1e6,1), {k:v['ms'] for k,v in d['kernels'].items()})"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gtconv_bwd_src -s 1 -c 1 -f -o $OUT/ncu_decoder_src_interleaved \
      python bench.py --workload decoder --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > $OUT/ncu_decoder_src.log 2>&1
timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log | cut -c1-200
timeout 900 python bench.py --workload model --steps 5 --warmup 2 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; tail -c 200 $OUT/bench_model.json
