"""Monte-Carlo simulator: the C++ body tracer against the Python one, the ring sampler, the NumPy
restatement of the simulation rules against the master equation (CPU), and the CUDA kernels against
the NumPy restatement bit for bit (GPU)."""

import numpy
import pytest
import scipy.integrate

from chemical_kinetics_and_program_execution_b200 import _lib, configs, programs
from oracle import mc_reference
import test_programs as tp
from test_random_programs import random_tree

BODIES = [(tp.decay, 'ex1-radioactive-decay', 2), (tp.ferromagnet, 'ex2-ferromagnetic-chain', 2),
          (tp.copolymerization, 'ex3-copolymerization', 4)]


@pytest.mark.parametrize('body,tag,size_a', BODIES)
def test_cpp_tracer_gives_the_tree_of_the_python_tracer(body, tag, size_a):
  """tapes_program_tree traces the compiled body of a reference problem; the Python restatement of
  the same body, traced by programs.trace, must give the same arrays."""
  mine, theirs = programs.trace(body, size_a), _lib.program_tree(tag)
  for key in mine:
    assert numpy.array_equal(mine[key], theirs[key]), key
  # a registered tree survives the round trip through the interpreter and the tracer
  _lib.register_program('mc-roundtrip', size_a, mine)
  again = _lib.program_tree('mc-roundtrip')
  for key in mine:
    assert numpy.array_equal(mine[key], again[key]), key


def test_ring_sampler_follows_the_table():
  size_a, cl_k, n = 3, 3, 300000
  table = configs.markov_table(size_a, cl_k, 7)
  tape = _lib.sample_ring(size_a, cl_k, table, n, 11)
  assert tape.dtype == numpy.uint8 and tape.size == n and tape.max() < size_a
  freq = mc_reference.window_counts(tape, size_a, cl_k) / n
  assert abs(freq - table).max() < 5 / numpy.sqrt(n)
  assert numpy.array_equal(tape, _lib.sample_ring(size_a, cl_k, table, n, 11))
  assert not numpy.array_equal(tape, _lib.sample_ring(size_a, cl_k, table, n, 12))


def master_equation(oracle, tag, size_a, cl_k, p0, t_end):
  f = oracle.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k, mode=oracle.MERGED)
  sol = scipy.integrate.solve_ivp(lambda t, y: f(y, t), (0, t_end), p0, method='DOP853', rtol=1e-9, atol=1e-12)
  return sol.y[:, -1]


def test_simulation_rules_reproduce_the_master_equation(oracle):
  """Radioactive decay has no correlations to neglect, so the closure is exact: the simulated
  window frequencies must follow the master equation within the statistical error."""
  size_a, cl_k, n, events = 2, 3, 40000, 40
  p0 = configs.markov_table(size_a, cl_k, 3)
  tree = programs.trace(tp.decay, size_a)
  tape = _lib.sample_ring(size_a, cl_k, p0, n, 5)
  for step in range(n // events):  # one unit of time
    tape = mc_reference.substep(tree, size_a, tape, events, 5, step)
  got = mc_reference.window_counts(tape, size_a, cl_k) / n
  want = master_equation(oracle, 'ex1-radioactive-decay', size_a, cl_k, p0, 1.0)
  assert abs(want - p0).max() > 0.3  # the table moved a lot
  assert abs(got - want).max() < 4 / numpy.sqrt(n)


@pytest.mark.gpu
def test_cuda_simulator_matches_numpy_rules_bit_for_bit():
  from chemical_kinetics_and_program_execution_b200 import markov_tapes as mt
  cases = [('ex1-radioactive-decay', programs.trace(tp.decay, 2), 2), ('ex2-ferromagnetic-chain', programs.trace(tp.ferromagnet, 2), 2),
           ('ex3-copolymerization', programs.trace(tp.copolymerization, 4), 4)]
  mt.register_program('mc-relay', 3, tp.relay)
  cases.append(('mc-relay', programs.trace(tp.relay, 3), 3))
  for seed in (103, 107):  # random programs: both heads, rewrites of rewritten cells, zero-weight options
    tree = random_tree(3, seed, 5)
    _lib.register_program(f'mc-rnd-{seed}', 3, tree)
    cases.append((f'mc-rnd-{seed}', _lib.program_tree(f'mc-rnd-{seed}'), 3))
  for tag, tree, size_a in cases:
    n, events, steps, seed = 5000, 700, 6, 9  # dense events: contested cells are common
    tape0 = _lib.sample_ring(size_a, 2, configs.markov_table(size_a, 2, 4), n, seed)
    want = tape0
    for step in range(steps):
      want = mc_reference.substep(tree, size_a, want, events, seed, step)
    ts = [steps * events / n]
    freq, got = mt.monte_carlo(tag=tag, size_a=size_a, cl_k=3, ts=ts, tape0=tape0, events_per_substep=events,
                               seed=seed, return_tape=True)
    assert numpy.array_equal(got, want), tag
    assert (want != tape0).any() or tag.startswith('mc-rnd'), tag
    assert numpy.array_equal(numpy.rint(freq[0] * n).astype(numpy.int64), mc_reference.window_counts(want, size_a, 3))


@pytest.mark.gpu
def test_cuda_simulator_follows_the_master_equation():
  """ex1 (exact closure) at 2^20 sites against the HBM-resident stepper; ex2 stays within a few
  statistical errors of its closure over a short time."""
  from chemical_kinetics_and_program_execution_b200 import markov_tapes as mt
  n = 1 << 20
  for tag, size_a, cl_k, p0, t_end, tol, moved in (
      ('ex1-radioactive-decay', 2, 4, configs.markov_table(2, 4, 3), 1.5, 5.0, 100.0),
      ('ex2-ferromagnetic-chain', 2, 5, configs.ex2_p0(5), 2.0, 8.0, 1.0)):  # the ferromagnet moves slowly
    ts = numpy.array([0.0, t_end / 2, t_end])
    sim = mt.monte_carlo(tag=tag, size_a=size_a, cl_k=cl_k, ts=ts, p0=p0, n_sites=n, seed=21)
    ode = mt.ode_integrate_device(tag=tag, size_a=size_a, cl_k=cl_k, p0=p0, ts=ts, rtol=1e-10, atol=1e-13)
    assert abs(sim.sum(axis=1) - 1).max() < 1e-12
    assert abs(sim[0] - p0).max() < 5 / numpy.sqrt(n)
    assert abs(sim - ode).max() < tol / numpy.sqrt(n), (tag, abs(sim - ode).max())
    assert abs(ode[-1] - ode[0]).max() > moved / numpy.sqrt(n)  # the comparison is not vacuous


def test_ferromagnet_chain_restatement_reproduces_the_reference_script():
  """the reference's own ferromagnet Monte Carlo
  (examples/ex2_ferromagnet_mc.py).  Golden: island counts produced by the script's `simulate` and
  `island_length_stats` on a small chain (tests/golden/make_golden_round2.py).  The inputs drawn
  like the script draws them plus the NumPy restatement of its update rule give the same counts,
  number for number - which pins the checker of the CUDA kernel."""
  import os
  from conftest import GOLDEN
  from make_golden_round2 import MC_SMALL as c
  from chemical_kinetics_and_program_execution_b200 import configs  # noqa: F401  (package import without the library)
  import importlib.util
  import sys
  # markov_tapes needs a GPU to import; the input generator is plain NumPy, so load the function alone
  inputs = _load_function('ferromagnet_mc_inputs')
  gold = numpy.load(os.path.join(GOLDEN, 'ex2_mc_chain_counts.npz'))['chain_counts']
  chain0, sites, uniforms, accept = inputs(n_trials=c['n_trials'], chain_length=c['chain_length'], n_steps=c['n_steps'],
                                           sites_per_pair=c['sites_per_pair'], trials_per_step=c['trials_per_step'],
                                           beta=c['beta'], J=c['J'], h=c['h'], seed_offset=c['seed_offset'])
  got = mc_reference.ferromagnet_chain_counts(chain0, sites, uniforms, accept)
  assert got.shape == gold.shape and numpy.array_equal(got, gold)
  assert gold[:, :, 1:].sum() > 1000  # the golden run is not trivial


def _load_function(name):
  """A top-level function of markov_tapes.py without importing the module (which initialises CUDA)."""
  import ast
  import os
  from conftest import ROOT
  path = os.path.join(ROOT, 'chemical_kinetics_and_program_execution_b200', 'markov_tapes.py')
  tree = ast.parse(open(path).read())
  ns = dict(numpy=numpy)
  for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name == name:
      exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
  return ns[name]


@pytest.mark.gpu
def test_cuda_ferromagnet_chains_reproduce_the_reference_script():
  """The CUDA kernel against the reference script's own counts (golden) and, on a chain of the
  script's size, against the NumPy restatement."""
  import os
  from conftest import GOLDEN
  from make_golden_round2 import MC_SMALL as c
  from chemical_kinetics_and_program_execution_b200 import markov_tapes as mt
  gold = numpy.load(os.path.join(GOLDEN, 'ex2_mc_chain_counts.npz'))['chain_counts']
  got = mt.ferromagnet_monte_carlo(n_trials=c['n_trials'], chain_length=c['chain_length'], n_steps=c['n_steps'],
                                   sites_per_pair=c['sites_per_pair'], trials_per_step=c['trials_per_step'], beta=c['beta'],
                                   J=c['J'], h=c['h'], seed_offset=c['seed_offset'], batch=2)
  assert numpy.array_equal(got, gold)
  # the script's chain length and trials per step (50 000 sites, 500 trials per step), fewer steps and trials
  kw = dict(n_trials=2, chain_length=50000, n_steps=40, sites_per_pair=250, trials_per_step=500)
  got = mt.ferromagnet_monte_carlo(**kw)
  want = mc_reference.ferromagnet_chain_counts(*mt.ferromagnet_mc_inputs(**kw))
  assert numpy.array_equal(got, want) and got[:, 0, 2].min() > 100  # pairs of up-spins to start with
