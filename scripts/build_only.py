"""Builds the structure of the synthetic bench workload and prints the build timings (GPU box).
usage: build_only.py A k R [repeats]"""
import os, sys, time
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
for i in range(reps):
  mt.u_lib.tapes_release_model(tag.encode(), k)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  m = device.DeviceModel(tag, k)
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  nodes = m.info['n_nodes'] + m.info['worlds_walked']
  expand_s = (m.timing['device_expand_ms'] + m.timing['host_enumerate_ms']) * 1e-3
  print(f'build {i}: wall {dt:.3f} s  timing={m.timing}  nodes={m.info["n_nodes"]} '
        f'states_expanded_per_s={nodes / expand_s:.3e} irregular_levels={m.info["irregular_levels"]}', flush=True)
