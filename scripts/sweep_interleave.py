"""Level-phase time of one right-hand side with the blocks of prefix groups evaluated in storage
order (seed-major) vs in prefix order across seeds (GPU box).  usage: sweep_interleave.py A k R"""
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
print({k_: m.info[k_] for k_ in ('n_states', 'n_nodes', 'nnz', 'n_levels', 'interleaved_levels')}, m.timing, flush=True)
ref = None
for flag in (0, 1, 0, 1):
  m.set_option('interleave_seeds', flag)
  ph = phases(m)
  o = out.cpu().numpy()
  ref = o if ref is None else ref
  print(f'interleave_seeds={flag}: levels_ms={ph[1]:.3f} flux_ms={ph[2]:.3f} marg_ms={ph[0]:.3f} '
        f'identical={numpy.array_equal(o, ref)}', flush=True)
