#!/bin/bash
# Round-end check in ONE gpurun call: scripts/gpu_check.sh (tests, smoke, bench, ncu) + the other workloads + one A/B of the row-block size.
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
bash scripts/gpu_check.sh $TAG
run() {
  local name=$1; shift; local w=$1; shift
  env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --workload $w --no-cpu-baseline --e2e-steps 1 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1]); k=d["kernels"]
    print("%-30s ms/step %.3f  fwd %.4f  bwd_dst %.4f  bwd_src %.4f  step_frac %.3f  [%s|%s|%s]"%("$name", d["ms_per_step"], k["fwd"]["ms"], k["bwd_dst"]["ms"], k["bwd_src"]["ms"], d["roofline_step"]["frac"], k["fwd"]["kernel"].split("<")[0][7:], k["bwd_dst"]["kernel"].split("<")[0][7:], k["bwd_src"]["kernel"].split("<")[0][7:]))
except Exception as ex: print("$name parse fail", ex, open("$OUT/$name.err").read()[-300:])
PY
}
run bench_processor processor X=1
run bench_decoder decoder X=1
run encoder_rb1 encoder AB2_TMA_RB=1
run encoder_rb3 encoder AB2_TMA_RB=3
run encoder_srctma encoder AB2_SRC_TMA_ALWAYS=1
for w in config1-enc config1-proc config1-dec; do
  timeout 200 python bench.py --steps 50 --warmup 5 --workload $w --no-cpu-baseline --e2e-steps 1 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python -c "import json;d=json.loads(open('$OUT/bench_$w.json').read().strip().splitlines()[-1]);print('$w', d['ms_per_step'], d['roofline_step']['frac'])"
done
