"""Ferromagnetic chain, Monte-Carlo benchmark: the experiment of the reference's
examples/ex2_ferromagnet_mc.py (100 trials on a 50 000-site ring, 4000 time steps of 500 single-spin
trials each, seeds 1000 + trial; 33-45, 167-191) run on the GPU with the script's own random numbers,
so `chain_counts` - stored in the script's cache file layout (ferromagnet_mc_chain_counts.npz, key
`chain_counts`, [trial, time step, island length]) - is what the script computes, number for number
(tests/test_monte_carlo.py).  Instead of the plot (194-219) the 10th / 50th / 90th percentiles are
printed beside the analytic approximation of examples/ex2_ferromagnet_analytic.py.

usage: python examples/ex2_ferromagnet_mc.py [n_trials]
"""
import os
import sys
import time

import _common  # noqa: F401
import numpy
import scipy.integrate

from chemical_kinetics_and_program_execution_b200 import markov_tapes as mt

NUM_TRIALS = int(sys.argv[1]) if len(sys.argv) > 1 else 100
CHAIN_LENGTH, NUM_TIME_STEPS, SITES_PER_PAIR = 50000, 4000, 250
NUM_TRIALS_PER_TIME_STEP = CHAIN_LENGTH // 100
beta, J, h = 1.0, 1.0, -0.25
t_max, t_steps = 40, 4000
DATA_FILE = 'ferromagnet_mc_chain_counts.npz'


def analytic():
  """Birth-death chain over island lengths (examples/ex2_ferromagnet_analytic.py:26-61)."""
  n = 20
  a, b = numpy.exp(-beta * 4 * J), numpy.exp(beta * 2 * h)
  m = numpy.zeros((n, n))
  m[0, 0] = -1
  for k in range(1, n):
    m[k - 1, k] += 2 * a
    m[k, k] -= 2 * a * (1 + b)
    m[k, k - 1] += 2 * a * b
  birth = numpy.zeros(n)
  birth[0] = numpy.exp(-8 * beta * J + 2 * beta * h)
  y0 = numpy.zeros(n)
  y0[1] = 1 / SITES_PER_PAIR
  ts = numpy.linspace(0, t_max, t_steps)
  return numpy.clip(scipy.integrate.odeint(lambda y, t: m @ y + birth, y0, ts, rtol=1e-10, atol=1e-10), 0, numpy.inf)


if not os.access(DATA_FILE, os.R_OK):
  t0 = time.perf_counter()
  chain_counts = mt.ferromagnet_monte_carlo(n_trials=NUM_TRIALS, chain_length=CHAIN_LENGTH, n_steps=NUM_TIME_STEPS,
                                            sites_per_pair=SITES_PER_PAIR, trials_per_step=NUM_TRIALS_PER_TIME_STEP,
                                            beta=beta, J=J, h=h, seed_offset=1000)
  print(f'{NUM_TRIALS} trials x {NUM_TIME_STEPS} time steps on {CHAIN_LENGTH} sites: {time.perf_counter() - t0:.1f} s '
        '(most of it drawing the random numbers with numpy.random.RandomState on the host)')
  numpy.savez_compressed(DATA_FILE, chain_counts=chain_counts)

chain_counts = numpy.load(DATA_FILE)['chain_counts']
p10, p50, p90 = (numpy.percentile(chain_counts, q, axis=0) / CHAIN_LENGTH for q in (10, 50, 90))
aa = analytic()
ts = numpy.linspace(0, t_max, t_steps)
print('island probabilities p(L): Monte Carlo 10th / 50th / 90th percentile over the trials, analytic approximation')
for length in (1, 2, 3, 4):
  for i in (500, 1000, 2000, 3999):
    print(f'  L = {length}  t = {ts[i]:5.1f}   MC {p10[i, length]:.3e} / {p50[i, length]:.3e} / {p90[i, length]:.3e}'
          f'   analytic {aa[i, length - 1]:.3e}')
