"""Round-2 golden fixtures.  Runs only where /root/reference exists (the build container); the GPU
box uses the committed files.

1. ex2_analytic.npz: the reference's analytic approximation of the ferromagnetic chain, produced
   by importing examples/ex2_ferromagnet_analytic.py itself (pure NumPy / SciPy) and calling
   `get_p_history(t_max=60)` (the call of examples/ex2_ferromagnet_tape.py:114).  Stored: the island
   probabilities for L = 1..5 at the 1001 output times.
2. ex3_k12_trajectory.npz: BASELINE.json config 3: the long-chain variant of
   examples/ex3_copolymerization.py (A = 4, cl_k = 12, 1.68e7 states), p0 = the vectorised equivalent of
   its `get_p0` (checked against the reference generator at cl_k = 6 by tests/test_oracle.py), CPU
   oracle (merged mode) through solve_ivp DOP853 at rtol = atol = 1e-10 over t in [0, 10]; stored: the
   observables of examples/ex3_copolymerization.py:112-118 at 11 output times, the number of
   right-hand sides, and the end state (sparse).

3. ex2_mc_chain_counts.npz: the island counts of the reference's ferromagnet Monte Carlo
   (examples/ex2_ferromagnet_mc.py) on a small chain, produced by the script's own functions.

Usage: python tests/golden/make_golden_round2.py [--mc-only]
"""

import importlib.util
import os
import sys
import time

import numpy
import scipy.integrate

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from chemical_kinetics_and_program_execution_b200 import configs  # noqa: E402

EX3_SEQS = [[0, 1, 0], [0, 2, 0], [0, 1, 2, 0], [0, 2, 1, 3, 0], [0, 2, 1, 2, 0], [1, 3, 1, 2], [1, 3, 1, 3]]


def make_ex2_analytic():
  spec = importlib.util.spec_from_file_location(
      'ex2_ferromagnet_analytic', os.path.join(REF, 'examples', 'ex2_ferromagnet_analytic.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  hist = mod.get_p_history(t_max=60)
  numpy.savez_compressed(os.path.join(HERE, 'ex2_analytic.npz'), ts=numpy.linspace(0, 60, 1001),
                         islands=hist[:, :5])
  print('wrote ex2_analytic.npz', hist.shape)


def seq_sum(y, size_a, cl_k, seq):
  return float(numpy.asarray(y).reshape([size_a] * cl_k)[(Ellipsis,) + tuple(seq)].sum())


def make_ex3_long_chain():
  cl_k, size_a = 12, 4
  p0 = configs.ex3_p0(cl_k)
  f = oracle.get_dy_dt(tag='ex3-copolymerization', size_a=size_a, cl_k=cl_k, mode=oracle.MERGED)
  ts = numpy.linspace(0, 10, 11)
  t0 = time.time()
  sol = scipy.integrate.solve_ivp(lambda t, y: f(y, t), (0.0, 10.0), p0, t_eval=ts, method='DOP853',
                                  rtol=1e-10, atol=1e-10)
  obs = numpy.array([[seq_sum(sol.y[:, i], size_a, cl_k, s) for s in EX3_SEQS] for i in range(ts.size)])
  end = sol.y[:, -1]
  idx = numpy.nonzero(end)[0]
  numpy.savez_compressed(os.path.join(HERE, 'ex3_k12_trajectory.npz'), ts=ts, observables=obs,
                         nfev=numpy.array([sol.nfev]), end_idx=idx.astype(numpy.int64), end_val=end[idx])
  print(f'wrote ex3_k12_trajectory.npz: nfev={sol.nfev} {time.time() - t0:.1f}s non-zeros at t=10: {idx.size}')




# ---- ferromagnet Monte Carlo ----------------------------------------
MC_SMALL = dict(n_trials=3, chain_length=2000, n_steps=300, sites_per_pair=40, trials_per_step=20, seed_offset=1000,
                beta=1.0, J=1.0, h=-0.25)


def make_ferromagnet_mc():
  """ex2_mc_chain_counts.npz: examples/ex2_ferromagnet_mc.py's own `simulate` (46-122) and
  `island_length_stats` (134-163), extracted from the script (which cannot be imported: it runs the
  full experiment and plots at import time) and run with its driver loop (169-191) on a small chain."""
  import ast
  path = os.path.join(REF, 'examples', 'ex2_ferromagnet_mc.py')
  tree = ast.parse(open(path).read())
  ns = dict(numpy=numpy)
  for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ('simulate', 'island_length_stats'):
      exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
  c = MC_SMALL
  counts = numpy.zeros([c['n_trials'], c['n_steps'], 6])
  for n_trial in range(c['n_trials']):
    rng = numpy.random.RandomState(seed=n_trial + c['seed_offset'])
    pair_positions = rng.uniform(0, 1, size=c['chain_length']) < 1 / c['sites_per_pair']
    chain0 = (pair_positions | numpy.roll(pair_positions, 1)).astype(numpy.int8)
    history = ns['simulate'](chain0, c['n_steps'], num_trials_per_time_step=c['trials_per_step'], J=c['J'], h=c['h'],
                             beta=c['beta'], rng=rng)
    for n_time, chain in enumerate(history):
      stats = ns['island_length_stats'](chain)
      for c_len in range(1, 6):
        counts[n_trial, n_time, c_len] = stats.get(c_len, 0)
  numpy.savez_compressed(os.path.join(HERE, 'ex2_mc_chain_counts.npz'), chain_counts=counts)
  print('wrote ex2_mc_chain_counts.npz', counts.shape, counts.sum(axis=(0, 1)))


if __name__ == '__main__':
  if '--mc-only' not in sys.argv:
    oracle.build()
    make_ex2_analytic()
    make_ex3_long_chain()
  make_ferromagnet_mc()
