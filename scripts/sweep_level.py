"""Phase times for the values of one model option (GPU box).
usage: sweep_level.py A k R option v1 v2 ...   e.g.  sweep_level.py 10 8 24 ratio_table 0 1 0 1"""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import bench
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
option = sys.argv[4]
values = [int(v) for v in sys.argv[5:]]
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
p = bench.device_product_table(A, k, 3, torch.device('cuda'))
out = torch.empty_like(p)
m = device.DeviceModel(tag, k)
ref = None
for v in values:
  m.set_option(option, v)
  for _ in range(3):
    m.rhs(p, out)
  torch.cuda.synchronize()
  ph = numpy.zeros(3)
  for _ in range(8):
    ph += m.rhs_profile(p, out)
  ph /= 8
  o = out.cpu().numpy()
  ref = o if ref is None else ref
  print(f'{option}={v}: marg_ms={ph[0]:.3f} levels_ms={ph[1]:.3f} flux_ms={ph[2]:.3f} step_ms={ph.sum():.3f} '
        f'identical={numpy.array_equal(o, ref)}', flush=True)
