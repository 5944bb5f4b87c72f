#!/bin/bash
# multi-GPU check: NCCL parity test of the dst-row-sharded blocks + weak-scaling bench at N ranks
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi.log
tail -12 $OUT/pytest_multi.log
for n in 1 $N; do
  if [ $n == 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 50 --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  fi
  echo "bench n=$n exit $?"; tail -3 $OUT/bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n$n.json").read().strip().splitlines()[-1])
    print("n=$n value %.1f M edges/s"%(d["value"]/1e6), "ms/step %.3f"%d["ms_per_step"], d["config"]["edges_total"], d["config"]["edges_rank0"], d["config"]["src_rows_rank0"], "e2e", d["e2e"].get("ms_per_step"), d["e2e"].get("value"))
except Exception as ex: print("parse fail", ex)
PY
done
