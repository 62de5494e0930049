"""Right-hand-side time of BASELINE.json's configurations on the GPU beside the CPU port (GPU box)."""
import os, sys, time
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt
from oracle import oracle

oracle.build()
rules = configs.autocatalysis_rule_set()
mt.register_rule_set('autocatalysis-tape', 4, rules)
oracle.register_rules('autocatalysis-tape', 4, rules)
cases = [('ex2-ferromagnetic-chain', 2, 7, configs.ex2_p0(7)),
         ('autocatalysis-tape', 4, 10, configs.product_table([0.5, 0.3, 0.1, 0.1], 10)),
         ('ex3-copolymerization', 4, 12, configs.dirichlet_product_table(4, 12, 3)),
         ('ex4-chemical-turing', 9, 5, configs.markov_table(9, 5, 2)),
         ('ex5-msrtf-machine', 5, 5, configs.markov_table(5, 5, 2))]
for tag, size_a, cl_k, p in cases:
  m = device.DeviceModel(tag, cl_k)
  d_p = torch.from_numpy(p).cuda()
  out = torch.empty_like(d_p)
  for _ in range(3): m.rhs(d_p, out)
  torch.cuda.synchronize()
  reps = 50
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): m.rhs(d_p, out)
  e1.record(); torch.cuda.synchronize()
  gpu_ms = e0.elapsed_time(e1) / reps
  t0 = time.perf_counter()
  want, counters = oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED, want_counters=True)
  cpu_merged = time.perf_counter() - t0
  cpu_literal = None
  if size_a ** (2 * cl_k - 3) <= 2e8:  # the literal recursion visits about A^(2k-3) windows per one-cell rule
    t0 = time.perf_counter()
    oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.LITERAL)
    cpu_literal = time.perf_counter() - t0
  err = abs(out.cpu().numpy() - want).max() / abs(want).max()
  info, timing = m.info, m.timing
  print(f'{tag} A={size_a} k={cl_k}: states {info["n_states"]}, forest nodes {info["n_nodes"]}, nnz {info["nnz"]}, '
        f'build {timing["device_expand_ms"] + timing["device_csr_ms"] + timing["device_slices_ms"]:.1f} ms, '
        f'GPU rhs {gpu_ms * 1e3:.1f} us, CPU port merged {cpu_merged * 1e3:.2f} ms'
        + (f', literal (reference order) {cpu_literal * 1e3:.2f} ms' if cpu_literal is not None else '')
        + f', max rel dev {err:.1e}, timing {({k_: round(v_, 1) for k_, v_ in timing.items()})}', flush=True)
