#!/bin/bash
# exchange rounds and product variant of the fused peer exchange on N GPUs: r02_sweep_exchange.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
for unroll in 3 4; do
for chunks in 1 2 4 8 16; do
  TAPES_FLUX_UNROLL=$unroll timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --chunks $chunks --e2e-steps 1 > gpurun_out/sweep_ex.json 2> gpurun_out/sweep_ex.err
  python - <<PY
import json
try:
  d = json.load(open('gpurun_out/sweep_ex.json'))
  print('N=$N flux_unroll=$unroll rounds=$chunks: ms', round(d['ms_per_step'], 3), 'rank_ms', round(max(d['rank_compute_ms']), 3), 'exposed', round(d['exchange_exposed_ms'], 3),
        'strong', d['strong_scaling'].get('ms_per_step'), d['strong_scaling'].get('speedup'))
except Exception as ex:
  print('N=$N unroll=$unroll rounds=$chunks failed', repr(ex))
PY
done
done
