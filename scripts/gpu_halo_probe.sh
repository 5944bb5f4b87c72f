#!/bin/bash
N=${1:-4}; TAG=${2:-probe}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1
for env in "X=1" "NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_NCHANNELS_PER_NET_PEER=8"; do
  echo "== env: $env"
  env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/halo_probe.py 2> $OUT/probe.err | tee -a $OUT/probe.log | cut -c1-900
  tail -2 $OUT/probe.err | cut -c1-200
done
