#!/bin/bash
N=${1:-4}; TAG=${2:-trace}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
AB2_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
      bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
grep "trace rank" $OUT/bench.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1]); print("ms/step", d["ms_per_step"])
PY
