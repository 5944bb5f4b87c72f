#!/bin/bash
# quick iteration: gpu tests + bench lines (encoder headline, decoder, processor); optional ncu of the three kernels
TAG=${1:-iter}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<PY
import json
for name in ("bench",):
    try:
        d=json.loads(open("$OUT/%s.json"%name).read().strip().splitlines()[-1])
        print(name, "value %.1f M edges/s"%(d["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], "step frac", d["roofline_step"]["frac"], d["kernels"], "e2e", d["e2e"].get("ms_per_step"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex: print("parse fail", ex)
PY
for w in decoder processor; do
  timeout 600 python bench.py --steps 50 --warmup 5 --workload $w --no-cpu-baseline --e2e-steps 2 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$w.json").read().strip().splitlines()[-1])
    print("$w", "value %.1f M edges/s"%(d["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], "step frac", d["roofline_step"]["frac"], d["kernels"])
except Exception as ex: print("parse fail", ex)
PY
done
for w in graphconv model; do
  timeout 900 python bench.py --steps 20 --warmup 3 --workload $w > $OUT/bench_$w.json 2> $OUT/bench_$w.err; echo "bench $w exit $?"; tail -3 $OUT/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$w.json").read().strip().splitlines()[-1])
    print("$w", d["metric"], d["value"], "ms/step %.3f"%d["ms_per_step"], d.get("roofline"), d.get("peak_mem_GB"))
except Exception as ex: print("parse fail", ex)
PY
done
if [ "${NCU:-0}" == "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gtconv -s 3 -c 3 -f -o $OUT/gtconv_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_full.log 2>&1
fi
