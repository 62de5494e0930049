"""MSRTF machine (examples/ex5_msrtf_machine.py:45-71 of the reference: k = 5, uniform table over
the symbols {0, 1, 2}, t = 0..500 in 4001 points, DOP853 at rtol = atol = 1e-13)."""
import _common  # noqa: F401
import time

import numpy

from chemical_kinetics_and_program_execution_b200 import configs, markov_tapes as mt

ts = numpy.linspace(0, 500, 4001)
symbols = [[s] for s in range(5)]
t0 = time.perf_counter()
series, stats = mt.ode_integrate_device(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=configs.ex5_p0(5), ts=ts,
                                        rtol=1e-13, atol=1e-13, observables=symbols, return_states=False,
                                        want_stats=True)
print(f'{time.perf_counter() - t0:.2f} s, {stats["nfev"]} right-hand sides, {stats["accepted"]} steps')
print('symbol frequencies')
for i in (0, 400, 800, 2000, 4000):
  print(f'  t = {ts[i]:6.1f}  ' + '  '.join(f'{v:.6f}' for v in series[i]) + f'   sum {series[i].sum():.12f}')
