#!/bin/bash
# A/B of kernel variants through env switches: AB2_TMA (bit0 fwd, bit1 bwd_dst), AB2_ROW_BLOCKS
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
run() {  # name, env..., workload
  local name=$1; shift
  local w=$1; shift
  env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --workload $w --no-cpu-baseline --e2e-steps 1 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("%-28s ms/step %.3f  fwd %.4f  bwd_dst %.4f  bwd_src %.4f  step_frac %.3f"%("$name", d["ms_per_step"], k["fwd"]["ms"], k["bwd_dst"]["ms"], k["bwd_src"]["ms"], d["roofline_step"]["frac"]))
except Exception as ex: print("$name parse fail", ex, open("$OUT/$name.err").read()[-300:])
PY
}
for w in encoder processor decoder; do
  run ${w}_tma11 $w AB2_TMA=11
  run ${w}_tma3 $w AB2_TMA=3
done
run encoder_tma15 encoder AB2_TMA=15
run encoder_tma11_b encoder AB2_TMA=11
run encoder_tma8 encoder AB2_TMA=8
