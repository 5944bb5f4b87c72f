#!/bin/bash
# A/B of the warp-cooperative LDG src pass (rows per group) against the per-thread kernel.
TAG=${1:-ab4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
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
run encoder_src_thread encoder AB2_SRC_WARP=0
run encoder_src_r8 encoder AB2_SRC_ROWS=8
run encoder_src_r12 encoder AB2_SRC_ROWS=12
run encoder_src_r24 encoder AB2_SRC_ROWS=24
run encoder_src_r31 encoder AB2_SRC_ROWS=31
run processor_ldg_warp processor AB2_TMA=0
run processor_ldg_thread processor AB2_TMA=0 AB2_SRC_WARP=0
run config1proc_warp config1-proc X=1
run config1proc_thread config1-proc AB2_SRC_WARP=0
run config1enc_warp config1-enc X=1
run config1enc_thread config1-enc AB2_SRC_WARP=0
