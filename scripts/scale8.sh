#!/bin/bash
# weak-scaling variants of bench.py on N GPUs (run on a --gpus N box)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
mkdir -p gpurun_out
for variant in "--exchange rs_ag" "--exchange allreduce" "--exchange allreduce --rules-per-gpu 16"; do
  name=$(echo "$variant" | tr -d ' -' )
  timeout 400 $TR bench.py --gpus $N $variant --steps 30 > gpurun_out/scale_n${N}_${name}.json 2> gpurun_out/scale_n${N}_${name}.err
  echo "variant [$variant] exit $?"
  tail -1 gpurun_out/scale_n${N}_${name}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], 'rank_ms', [round(x,2) for x in d['rank_compute_ms']], 'exchange_ms', d['exchange_ms'], d['clocks'])"
done
