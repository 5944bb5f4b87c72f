#!/bin/bash
N=${1:-4}; TAG=${2:-scaleab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
i=0
for env in "AB2_PUSH_CTAS=64" "AB2_PUSH_CTAS=16" "AB2_PUSH_CTAS=1184" "AB2_OVERLAP=0 AB2_PUSH_CTAS=64" "AB2_HALO=nccl"; do
  i=$((i+1))
  env $env timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$i \
      bench.py --gpus $N --steps 50 --warmup 5 > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$i.json").read().strip().splitlines()[-1])
    print("%-40s value %.1f M edges/s  ms/step %.3f"%("$env", d["value"]/1e6, d["ms_per_step"]))
except Exception as ex: print("$env parse fail", ex, open("$OUT/bench_$i.err").read()[-400:])
PY
done
