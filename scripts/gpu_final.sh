#!/bin/bash
# Round-end check in ONE gpurun call: scripts/gpu_check.sh (tests, smoke, bench, ncu) + the other workloads + bandwidth probe.
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
run bench_config1-enc config1-enc X=1
run bench_config1-proc config1-proc X=1
run bench_config1-dec config1-dec X=1
run config1-enc_r4 config1-enc AB2_SRC_ROWS=4
run config1-proc_r4 config1-proc AB2_SRC_ROWS=4
run config1-dec_r4 config1-dec AB2_SRC_ROWS=4
timeout 120 python scripts/hbm_probe.py > $OUT/hbm_probe.json 2>&1; cat $OUT/hbm_probe.json
timeout 300 python bench.py --workload graphconv --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_graphconv.json 2> $OUT/bench_graphconv.err; tail -c 600 $OUT/bench_graphconv.json
