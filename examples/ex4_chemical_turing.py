"""Chemical Turing machine: fuel (P) drives a cursor that writes the bit string I O I.

Parameters of the reference's examples/ex4_chemical_turing.py:86-115 (k = 5, powered fractions
0.04 and 0.01, t = 0..2000 in 2001 points, DOP853 at rtol = atol = 1e-13); the end points are
compared with the values recorded there (lines 150-170).  The table stays in HBM; only the eight
sequence probabilities the reference plots cross the host boundary.
"""
import _common  # noqa: F401
import json
import os
import time

import numpy

from chemical_kinetics_and_program_execution_b200 import configs, markov_tapes as mt

golden = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'reference_known_answers.json')))
names = 'OAOOO OIBOO OIBIO OIOCO OIOCI OIOID P X'.split()
ts = numpy.linspace(0, 2000, 2001)
for label, fraction in (('a', 0.04), ('b', 0.01)):
  p0 = configs.ex4_p0(5, powered_fraction=fraction)
  t0 = time.perf_counter()
  series, stats = mt.ode_integrate_device(tag='ex4-chemical-turing', size_a=9, cl_k=5, p0=p0, ts=ts, rtol=1e-13,
                                          atol=1e-13, observables=golden['ex4_observables'], return_states=False,
                                          want_stats=True)
  print(f'powered fraction {fraction}: {time.perf_counter() - t0:.2f} s, {stats["nfev"]} right-hand sides')
  for name, got, want in zip(names, series[-1], golden[f'ex4_p0_{label}_t2000']):
    print(f'  p({name:5s}) at t = 2000: {got:.15e}   reference: {want:.15e}   rel. dev. {abs(got - want) / want:.1e}')
