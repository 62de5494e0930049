#!/bin/bash
# weak and strong scaling bench lines on N GPUs (run on a --gpus N box): r02_multi_gpu_bench.sh N
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
mkdir -p gpurun_out
for mode in weak strong; do
  timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --scaling $mode > gpurun_out/r02_h_bench_n${N}_${mode}.json 2> gpurun_out/r02_h_bench_n${N}_${mode}.err
  echo "$mode rc=$?"
  python - <<PY
import json
try:
  d = json.load(open('gpurun_out/r02_h_bench_n${N}_${mode}.json'))
  print('${mode}', 'N', d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'value', round(d['value'], 1), 'rank_ms', [round(x, 2) for x in d['rank_compute_ms']],
        'exposed', round(d['exchange_exposed_ms'], 3), 'e2e_ms', d['e2e'].get('ms_per_step'), 'parity', d['parity_check']['ok'], d['clocks'])
except Exception as ex:
  print('${mode}', repr(ex))
PY
done
