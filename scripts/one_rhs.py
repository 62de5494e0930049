"""Builds the synthetic bench structure and evaluates a few right-hand sides (for ncu captures).
usage: one_rhs.py A k R [n_rhs]"""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n_rhs = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
m = device.DeviceModel(tag, k)
p = bench.device_product_table(A, k, 3, torch.device('cuda'))
out = torch.empty_like(p)
for _ in range(n_rhs):
  m.rhs(p, out)
torch.cuda.synchronize()
print(m.info, flush=True)
