#!/bin/bash
TAG=${1:-ab3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
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
run encoder_default encoder X=1
run processor_default processor X=1
run processor_rb1 processor AB2_TMA_RB=1
run processor_rb2 processor AB2_TMA_RB=2
run decoder_default decoder X=1
run decoder_rb1 decoder AB2_TMA_RB=1
run decoder_rb3 decoder AB2_TMA_RB=3
run decoder_rb6 decoder AB2_TMA_RB=6
