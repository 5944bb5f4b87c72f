#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list and full captures of the three conv kernels.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launch.log 2>&1
grep -c gtconv $OUT/launches.csv
echo "== ncu full (fwd, bwd_dst, bwd_src)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gtconv -s 3 -c 3 -f -o $OUT/gtconv_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ls -la $OUT
fi
if [ "${SANITIZE:-0}" == "1" ]; then
echo "== compute-sanitizer memcheck (golden cases)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gtconv.py tests/test_gpu_csr.py -x -q -m gpu -k "golden or long_segments or empty" > $OUT/sanitizer.log 2>&1
echo "sanitizer exit $?" | tee -a $OUT/sanitizer.log; tail -5 $OUT/sanitizer.log
fi
