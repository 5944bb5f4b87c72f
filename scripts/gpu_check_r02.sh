#!/bin/bash
# One gpurun call: full GPU test suite, smoke, bench (headline), extra report lines.
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -40 $OUT/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -c 1500 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
tail -c 800 $OUT/bench_reference.json
fi
if [ "${MODEL:-0}" == "1" ]; then
echo "== model"
timeout 900 python bench.py --workload model --steps 5 --warmup 2 --profile > $OUT/bench_model.json 2> $OUT/bench_model.err; echo "model exit $?"
tail -c 3000 $OUT/bench_model.json; tail -5 $OUT/bench_model.err
AB2_TC=0 timeout 900 python bench.py --workload model --steps 5 --warmup 2 > $OUT/bench_model_tc0.json 2> $OUT/bench_model_tc0.err
tail -c 600 $OUT/bench_model_tc0.json
fi
