#!/bin/bash
TAG=${1:-lean}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for n in ${SCALE_NS:-8 4}; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n \
      bench.py --gpus $n --steps 50 --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n$n.json").read().strip().splitlines()[-1])
    print("n=$n value %.1f M edges/s  ms/step %.3f halo %s"%(d["value"]/1e6, d["ms_per_step"], d["config"].get("halo_rows_rank0")))
except Exception as ex: print("n=$n parse fail", ex, open("$OUT/bench_n$n.err").read()[-400:])
PY
done
