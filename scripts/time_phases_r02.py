"""One GPU, bench structure: device time of the weights phase alone, the product alone and the whole
right-hand side, each as back-to-back calls between two CUDA events, beside the per-phase figures of
rhs_profile.  usage: time_phases_r02.py [A k R]"""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (10, 8, 24)
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
m = device.DeviceModel(tag, k)
p = bench.device_product_table(A, k, 3, torch.device('cuda'))
out = torch.empty_like(p)
n = A ** k


def timed(f, reps=20):
  for _ in range(5):
    f()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    f()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps


print(f'weights alone          {timed(lambda: m.weights(p)):.3f} ms')
print(f'product alone          {timed(lambda: m.flux_rows(out, 0, n)):.3f} ms')
print(f'weights + product      {timed(lambda: (m.weights(p), m.flux_rows(out, 0, n))):.3f} ms')
print(f'right-hand side        {timed(lambda: m.rhs(p, out)):.3f} ms')
print('rhs_profile phases    ', [round(float(x), 3) for x in m.rhs_profile(p, out)])
