"""One problem on several GPUs of one node (one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/multi_gpu.py

The library deals the flux rules of the problem to the ranks (tapes_model_part); every rank evaluates
its share over the whole table and the partial dy/dt are summed inside the product kernel over NVLink
peer memory (parallel.PeerExchangeRhs).  The stepper then runs on all ranks in lockstep and every rank
ends up with the same bits.  The calls are those of scripts/check_multi_gpu.py, which also compares
them with one GPU evaluating the whole problem.
"""
import _common  # noqa: F401
import os

import numpy
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))  # carries the IPC handles and the barriers only

from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt, parallel  # noqa: E402

tag, size_a, cl_k = 'ex3-copolymerization', 4, 10                     # 1 048 576 states, 8 flux rules
share = device.DeviceModel(tag, cl_k, part=(rank, world))
peer = parallel.PeerExchangeRhs(share)
print(f'rank {rank}: {share.info["n_flux_rules"]} flux rules, {share.info["n_nodes"]} forest nodes', flush=True)

ts = numpy.linspace(0.0, 10.0, 11)
series = mt.ode_integrate_device(tag=tag, size_a=size_a, cl_k=cl_k, p0=configs.ex3_p0(cl_k), ts=ts, rtol=1e-10,
                                 atol=1e-12, observables=[[1], [2], [3], [1, 2], [1, 3]], return_states=False,
                                 peer_group=peer)
if rank == 0:
  print('t, p(A), p(M), p(N), p(AM), p(AN)')
  for t, row in zip(ts, series):
    print(f'{t:5.1f}  ' + '  '.join(f'{v:.6e}' for v in row))
peer.close()
dist.destroy_process_group()
