#!/bin/bash
OUT=gpurun_out/${1:-r02g}
mkdir -p $OUT
timeout 600 python bench.py --workload graphconv --steps 10 --warmup 3 --profile > $OUT/bench_graphconv.json 2> $OUT/bench_graphconv.err; tail -c 300 $OUT/bench_graphconv.err
AB2_TC=0 timeout 600 python bench.py --workload graphconv --steps 10 --warmup 3 > $OUT/bench_graphconv_tc0.json 2> $OUT/bench_graphconv_tc0.err
timeout 600 python bench.py --workload edgepath --steps 10 --warmup 3 > $OUT/bench_edgepath.json 2> $OUT/bench_edgepath.err; tail -c 1500 $OUT/bench_edgepath.json; tail -c 300 $OUT/bench_edgepath.err
for w in decoder processor; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
timeout 600 python -m pytest tests/test_gpu_graphconv_blocks.py tests/test_gpu_parity_r02.py -q -m gpu -k "graphconv" > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
