"""A problem of one's own, written as a Python function over the reference's three primitives
(tape-get, tape-set!, choose; framework/gambit_macros.scm:99-125) - no Scheme, no rebuild.

A program-tape token T next to a data-tape pair (x, y) with x != y copies x over y with a rate that
depends on x; used tokens become blanks.  The master equation is integrated on the GPU and checked
against a Monte-Carlo simulation of the same program."""
import _common  # noqa: F401
import numpy

from chemical_kinetics_and_program_execution_b200 import configs, markov_tapes as mt

BLANK, TOKEN = 0, 2


def copier(tape):
  if tape.get(False, 0) != TOKEN:            # (tape-get #f 0): program tape under the head
    return
  x, y = tape.get(True, 0), tape.get(True, 1)
  if x == y:
    return
  if tape.choose_value([(0.2 + 0.3 * x, True), (0.8 - 0.3 * x, False)]):
    tape.set(True, 1, x)                     # (tape-set! #t 1 x)
    tape.set(False, 0, BLANK)


mt.register_program('copier', 3, copier)
cl_k, ts = 4, numpy.linspace(0, 6, 7)
p0 = configs.product_table([0.5, 0.2, 0.3], cl_k)
ode = mt.ode_integrate_device(tag='copier', size_a=3, cl_k=cl_k, p0=p0, ts=ts, rtol=1e-10, atol=1e-12)
n_sites = 1 << 22
sim = mt.monte_carlo(tag='copier', size_a=3, cl_k=cl_k, ts=ts, p0=p0, n_sites=n_sites, seed=3)
table = lambda y: y.reshape((len(ts),) + (3,) * cl_k)
print(f'symbol frequencies: master equation | Monte Carlo on {n_sites} sites')
for i, t in enumerate(ts):
  a = [mt.seq_prob(table(ode), (s,), num_prefix_indices=1)[0][i] for s in range(3)]
  b = [mt.seq_prob(table(sim), (s,), num_prefix_indices=1)[0][i] for s in range(3)]
  print(f'  t = {t:3.0f}   ' + ' '.join(f'{v:.5f}' for v in a) + '  |  ' + ' '.join(f'{v:.5f}' for v in b))
print(f'largest difference over all {3 ** cl_k} table entries and times: {abs(ode - sim).max():.1e} '
      f'(statistical error ~{n_sites ** -0.5:.0e}; the rest is what the closure neglects)')
