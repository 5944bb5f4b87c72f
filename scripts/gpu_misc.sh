#!/bin/bash
TAG=${1:-misc}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_gtconv.py -x -q -m gpu -k "config1 or golden" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest.log
for w in config1-enc config1-proc config1-dec; do
  timeout 300 python bench.py --steps 200 --warmup 20 --workload $w --no-cpu-baseline --e2e-steps 2 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$w.json").read().strip().splitlines()[-1]); k=d["kernels"]
    print("$w value %.1f M edges/s ms/step %.4f"%(d["value"]/1e6, d["ms_per_step"]), {x:(k[x]["kernel"].split("<")[0], k[x]["ms"]) for x in k}, "step_frac", d["roofline_step"]["frac"], "e2e", d["e2e"]["ms_per_step"])
except Exception as ex: print("$w parse fail", ex, open("$OUT/bench_$w.err").read()[-300:])
PY
done
timeout 300 python bench.py --steps 20 --warmup 3 --workload graphconv --profile > $OUT/bench_graphconv.json 2> $OUT/bench_graphconv.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_graphconv.json").read().strip().splitlines()[-1])
print("graphconv ms/step", d["ms_per_step"], d["roofline"]["achieved"])
for r in d["kernel_breakdown"]: print(r)
PY
