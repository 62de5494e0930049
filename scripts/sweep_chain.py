"""Level-phase time of one right-hand side with the general level kernel vs the lean chain kernels
and their batchings (GPU box).  usage: sweep_chain.py A k R"""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import bench
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
p = bench.device_product_table(A, k, 3, torch.device('cuda'))
out = torch.empty_like(p)


def phases(m, reps=8):
  for _ in range(3):
    m.rhs(p, out)
  torch.cuda.synchronize()
  ph = numpy.zeros(3)
  for _ in range(reps):
    ph += m.rhs_profile(p, out)
  return ph / reps


m = device.DeviceModel(tag, k)
print({k_: m.info[k_] for k_ in ('n_states', 'n_nodes', 'nnz', 'n_levels', 'chain_levels', 'launches_per_rhs')},
      m.timing, flush=True)
m.set_option('chain_kernels', 0)
ph = phases(m)
ref = out.cpu().numpy()
print(f'general level kernel: levels_ms={ph[1]:.3f} flux_ms={ph[2]:.3f} marg_ms={ph[0]:.3f}', flush=True)
m.set_option('chain_kernels', 1)
for unroll in (62, 63, 61, 64, 52, 53, 54, 55, 45, 81, 82, 62):
  m.set_option('chain_unroll', unroll)
  ph = phases(m)
  same = numpy.array_equal(out.cpu().numpy(), ref)
  print(f'chain kernels blocks/SM={unroll // 10} parents in flight={unroll % 10}: levels_ms={ph[1]:.3f} '
        f'flux_ms={ph[2]:.3f} identical={same}', flush=True)
