#!/bin/bash
# weak-scaling run like the driver's: N = 1, 2, 4, 8 back to back on one box (+ NCCL parity test)
TAG=${1:-scale}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_multi.log
for n in ${SCALE_NS:-1 2 4 8}; do
  if [ $n == 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
        bench.py --gpus $n --steps 50 --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  fi
  echo "bench n=$n exit $?"; tail -2 $OUT/bench_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n$n.json").read().strip().splitlines()[-1])
    c=d["config"]
    print("n=$n value %.1f M edges/s"%(d["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], "edges_total", c["edges_total"], "rank0 edges", c["edges_rank0"], "src", c["src_rows_rank0"], "halo", c.get("halo_rows_rank0"))
except Exception as ex: print("parse fail", ex)
PY
done
