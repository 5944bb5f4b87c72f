#!/bin/bash
OUT=gpurun_out/${1:-r02k}
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log | cut -c1-300
for gph in encoder decoder processor; do
timeout 600 python bench.py --workload edgepath --edgepath-graph $gph --steps 10 --warmup 3 > $OUT/bench_edgepath_$gph.json 2> $OUT/bench_edgepath_$gph.err
python -c "
import json,sys
d=json.loads(open('$OUT/bench_edgepath_$gph.json').read().strip().splitlines()[-1]); print('$gph', {k:round(v,3) for k,v in d.items() if k.startswith('ms_')})" 2>&1 | tail -1
done
timeout 900 python bench.py --workload model --steps 5 --warmup 2 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; echo "model exit $?"
tail -c 300 $OUT/bench_model.json; tail -3 $OUT/bench_model.err | cut -c1-300
timeout 600 python bench.py --workload graphconv --steps 10 --warmup 3 --profile > $OUT/bench_graphconv.json 2> $OUT/bench_graphconv.err
timeout 900 python bench.py --steps 20 --warmup 5 --config5 off > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 600 $OUT/bench.json
