"""Generates the committed golden fixtures.  Runs only where /root/reference exists (the build
container); the GPU box uses the committed .npz/.json files.

1. p0 fixtures: the `get_p0` generators are extracted (ast) from the reference's example scripts
   and executed as they are; the non-zero entries of what they return are stored.  The scripts
   themselves cannot be imported (they import matplotlib and the Gambit-backed markov_tapes).
2. Reference known answers: the canary vector (framework/markov_tapes.py:362) and the 16 ex4
   end-point probabilities (examples/ex4_chemical_turing.py:150-170), transcribed into
   reference_known_answers.json.
3. Oracle trajectories: the CPU oracle (oracle/tape_oracle.cpp) driven by the same SciPy calls
   as the reference (framework/markov_tapes.py:318, 349-354) for ex2, ex4 and ex5; end points and
   a few observables are stored.  For ex4 the result is compared with (2) - this is what pins
   the oracle to the reference beyond the canary.

Usage: python tests/golden/make_golden.py [--skip-trajectories]
"""

import ast
import itertools
import json
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


def extract_functions(path, names, extra_globals):
  """Executes the named top-level function definitions (and simple constant assignments they
  need) of a reference script in an isolated namespace."""
  tree = ast.parse(open(path).read())
  ns = dict(numpy=numpy, itertools=itertools)
  ns.update(extra_globals)
  for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in names:
      exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
  return ns


def sparse(p):
  p = numpy.asarray(p, dtype=numpy.float64).ravel()
  idx = numpy.nonzero(p)[0]
  return idx.astype(numpy.int64), p[idx]


def make_p0_fixtures():
  out = {}
  ex = os.path.join(REF, 'examples')
  ns = extract_functions(os.path.join(ex, 'ex2_ferromagnet_tape.py'), {'get_p0'}, {})
  for k in range(3, 8):
    i, v = sparse(ns['get_p0'](k, p_pair=1 / 250))
    out[f'ex2_k{k}_idx'], out[f'ex2_k{k}_val'] = i, v
  ns = extract_functions(os.path.join(ex, 'ex3_copolymerization.py'), {'get_p0'}, dict(CL_K=6))
  i, v = sparse(ns['get_p0']())
  out['ex3_k6_idx'], out['ex3_k6_val'] = i, v
  ns = extract_functions(os.path.join(ex, 'ex4_chemical_turing.py'), {'get_p0'},
                         dict(SIZE_A=9, CL_K=5))
  for name, pf in (('a', 0.04), ('b', 0.01)):
    i, v = sparse(ns['get_p0'](powered_fraction=pf))
    out[f'ex4_{name}_idx'], out[f'ex4_{name}_val'] = i, v
  ns = extract_functions(os.path.join(ex, 'ex4var2_chemical_turing.py'), {'get_p0e'},
                         dict(SIZE_A=10, CL_K=5))
  i, v = sparse(ns['get_p0e']())
  out['ex4var2_idx'], out['ex4var2_val'] = i, v
  ns = extract_functions(os.path.join(ex, 'ex5_msrtf_machine.py'), {'get_p0'},
                         dict(SIZE_A=5, CL_K=5))
  i, v = sparse(ns['get_p0']())
  out['ex5_idx'], out['ex5_val'] = i, v
  numpy.savez_compressed(os.path.join(HERE, 'p0_fixtures.npz'), **out)
  print('wrote p0_fixtures.npz:', sorted(out))


def dense(idx, val, n):
  p = numpy.zeros(n)
  p[idx] = val
  return p


EX4_SEQS = [[5, 0, 5, 5, 5], [5, 4, 1, 5, 5], [5, 4, 1, 4, 5], [5, 4, 5, 2, 5], [5, 4, 5, 2, 4],
            [5, 4, 5, 4, 3], [6], [7]]  # examples/ex4_chemical_turing.py:131-140


EX2_FIXED_STEP = dict(method='DOP853', rtol=1e-3, atol=1e-6, max_step=0.05, first_step=0.05)


def observe(y, size_a, cl_k, seq):
  spd = numpy.asarray(y).reshape([size_a] * cl_k)
  picked = spd[(Ellipsis,) + tuple(seq)]
  return float(picked.sum())


def make_trajectories():
  fx = numpy.load(os.path.join(HERE, 'p0_fixtures.npz'))
  ka = json.load(open(os.path.join(HERE, 'reference_known_answers.json')))
  result = {}
  # ex4: DOP853 rtol=atol=1e-13, ts=linspace(0, 2000, 2001)
  for name in ('a', 'b'):
    p0 = dense(fx[f'ex4_{name}_idx'], fx[f'ex4_{name}_val'], 9 ** 5)
    for mode, mname in ((oracle.MERGED, 'merged'), (oracle.LITERAL, 'literal')):
      f = oracle.get_dy_dt(tag='ex4-chemical-turing', size_a=9, cl_k=5, mode=mode)
      t0 = time.time()
      sol = scipy.integrate.solve_ivp(lambda t, y: f(y, t), (0.0, 2000.0), p0,
                                      t_eval=numpy.linspace(0, 2000.0, 2001),
                                      rtol=1e-13, atol=1e-13, method='DOP853')
      y_end = sol.y[:, -1]
      obs = [observe(y_end, 9, 5, s) for s in EX4_SEQS]
      want = ka[f'ex4_p0_{name}_t2000']
      rel = [abs(o - w) / abs(w) for o, w in zip(obs, want)]
      print(f'ex4 p0_{name} {mname}: nfev={sol.nfev} {time.time() - t0:.1f}s max rel dev vs '
            f'reference golden = {max(rel):.3e}')
      result[f'ex4_{name}_{mname}_obs'] = numpy.array(obs)
      result[f'ex4_{name}_{mname}_nfev'] = numpy.array([sol.nfev])
      if mname == 'merged':
        i, v = sparse(y_end)
        result[f'ex4_{name}_end_idx'], result[f'ex4_{name}_end_val'] = i, v
        # mid-trajectory states for RHS parity tests at non-trivial support
        for tt in (10, 100, 1000):
          i, v = sparse(sol.y[:, tt])
          result[f'ex4_{name}_t{tt}_idx'], result[f'ex4_{name}_t{tt}_val'] = i, v
  # ex5: DOP853 1e-13, ts=linspace(0, 500, 4001)
  p0 = dense(fx['ex5_idx'], fx['ex5_val'], 5 ** 5)
  f = oracle.get_dy_dt(tag='ex5-msrtf-machine', size_a=5, cl_k=5, mode=oracle.MERGED)
  t0 = time.time()
  sol = scipy.integrate.solve_ivp(lambda t, y: f(y, t), (0.0, 500.0), p0,
                                  t_eval=numpy.linspace(0, 500.0, 4001),
                                  rtol=1e-13, atol=1e-13, method='DOP853')
  print(f'ex5: nfev={sol.nfev} {time.time() - t0:.1f}s')
  result['ex5_end'] = sol.y[:, -1]
  result['ex5_t50'] = sol.y[:, 400]
  result['ex5_nfev'] = numpy.array([sol.nfev])
  # ex2: odeint rtol=atol=1e-9, ts=linspace(0, 60, 1001), k=3..7
  for k in range(3, 8):
    p0 = dense(fx[f'ex2_k{k}_idx'], fx[f'ex2_k{k}_val'], 2 ** k)
    f = oracle.get_dy_dt(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=k, mode=oracle.LITERAL)
    ys = scipy.integrate.odeint(f, p0, numpy.linspace(0, 60, 1001), rtol=1e-9, atol=1e-9)
    result[f'ex2_k{k}_end'] = ys[-1]
    result[f'ex2_k{k}_t6'] = ys[100]
    # The same problem through DOP853 on a fixed step sequence (max_step forces every step to
    # h = 0.05 and the loose tolerance accepts all of them): with identical steps on both sides
    # the trajectories differ only through dy/dt rounding, which is what the 1e-12 parity bar is
    # about.  With adaptive steps a 1-ulp change of p0 already moves the k=7 end point by 8e-13
    # (one flipped step rejection), so adaptive runs cannot be compared that tightly.
    sol = scipy.integrate.solve_ivp(lambda t, y: f(y, t), (0.0, 60.0), p0, t_eval=[0.0, 30.0, 60.0],
                                    **EX2_FIXED_STEP)
    result[f'ex2_k{k}_dop853_end'] = sol.y[:, -1]
    result[f'ex2_k{k}_dop853_nfev'] = numpy.array([sol.nfev])
  numpy.savez_compressed(os.path.join(HERE, 'oracle_trajectories.npz'), **result)
  print('wrote oracle_trajectories.npz')


if __name__ == '__main__':
  make_p0_fixtures()
  if '--skip-trajectories' not in sys.argv:
    make_trajectories()
