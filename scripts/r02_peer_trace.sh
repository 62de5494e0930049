#!/bin/bash
# Time line of the fused exchange (TAPES_PEER_TRACE=1, csrc/flux.cu peer_rhs) on N GPUs: r02_peer_trace.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for chunks in ${2:-1 4 8}; do
  echo "== rounds $chunks"
  TAPES_PEER_TRACE=1 timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --chunks $chunks 2>&1 | grep "peer trace" | tail -2
done
