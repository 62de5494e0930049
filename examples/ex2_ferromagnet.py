"""Ferromagnetic chain: islands of up-spins in a down-magnetised chain melt and regrow.

Parameters of the reference's examples/ex2_ferromagnet_tape.py:74-84 (context lengths 3..7,
p_pair = 1/250, LSODA at rtol = atol = 1e-9, t = 0..60) with the analytic approximation of
examples/ex2_ferromagnet_analytic.py:26-61 and a Monte-Carlo run of the same program
(the reference's ex2_ferromagnet_mc.py simulates this one problem by hand).
"""
import _common  # noqa: F401
import numpy
import scipy.integrate

from chemical_kinetics_and_program_execution_b200 import configs, markov_tapes as mt

T_MAX, LENGTHS = 60.0, (1, 2, 3, 4, 5)
ts = numpy.linspace(0, T_MAX, 1001)
show = [0, 250, 500, 1000]


def islands(history, cl_k):
  """p(0 1^L 0) over time for the island lengths that fit the table."""
  table = history.reshape((len(history),) + (2,) * cl_k)
  return {L: mt.seq_prob(table, (0,) + (1,) * L + (0,), num_prefix_indices=1)[0] for L in LENGTHS}


def analytic(beta=1.0, J=1.0, h=-0.25, n_lengths=20, p_pair=1 / 250):
  """Birth-death chain over island lengths: melting at either end with rate a = exp(-4 beta J),
  growth with a * exp(2 beta h), single up-spins vanish at rate 1 and appear spontaneously."""
  a, b = numpy.exp(-4 * beta * J), numpy.exp(2 * beta * h)
  m = numpy.zeros((n_lengths, n_lengths))
  m[0, 0] = -1
  for k in range(1, n_lengths):
    m[k - 1, k] += 2 * a
    m[k, k] -= 2 * a * (1 + b)
    m[k, k - 1] += 2 * a * b
  birth = numpy.zeros(n_lengths)
  birth[0] = numpy.exp(-8 * beta * J + 2 * beta * h)
  y0 = numpy.zeros(n_lengths)
  y0[1] = p_pair
  return numpy.clip(scipy.integrate.odeint(lambda y, t: m @ y + birth, y0, ts, rtol=1e-10, atol=1e-10), 0, None)


by_k = {}
for cl_k in range(3, 8):
  ys = mt.ode_integrate(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=cl_k, p0=configs.ex2_p0(cl_k, p_pair=1 / 250),
                        ts=ts, odeint_kwargs=dict(rtol=1e-9, atol=1e-9))
  by_k[cl_k] = islands(ys, cl_k)
aa = analytic()
n_sites = 1 << 22
mc = mt.monte_carlo(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=7, ts=ts[show], p0=configs.ex2_p0(7, p_pair=1 / 250),
                    n_sites=n_sites, seed=1)
mc_islands = islands(mc, 7)
print(f'island probabilities p(0 1^L 0); Monte Carlo on {n_sites} sites (statistical error ~{n_sites ** -0.5:.0e})')
for L in LENGTHS:
  print(f'L = {L}')
  for j, i in enumerate(show):
    cols = '  '.join(f'k={k}: {by_k[k][L][i]:.3e}' for k in (3, 5, 7))
    print(f'  t = {ts[i]:5.1f}   {cols}   analytic: {aa[i, L - 1]:.3e}   Monte Carlo: {mc_islands[L][j]:.3e}')
