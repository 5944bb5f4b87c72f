#!/bin/bash
TAG=${1:-model}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_vs_unfused.py -x -q -m gpu -s > $OUT/pytest_compare.log 2>&1; echo "compare exit $?"; grep compare $OUT/pytest_compare.log; tail -3 $OUT/pytest_compare.log
timeout 900 python bench.py --steps 5 --warmup 3 --workload model --profile > $OUT/bench_model_profile.json 2> $OUT/bench_model_profile.err; echo "model exit $?"; tail -3 $OUT/bench_model_profile.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_model_profile.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "mem", d["peak_mem_GB"])
for r in d["kernel_breakdown"]: print(r)
PY
