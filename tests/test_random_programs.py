"""Randomly generated programs (decision trees over read / write / choose): the product's front end
and CUDA path against the oracle interpreting the same tree.  Seeds are fixed; the generator covers
what the shipped problems do not: reads far from the head on both tapes, repeated reads and writes
of a cell, zero-weight options, programs that write without ever changing anything."""

import numpy
import pytest

from chemical_kinetics_and_program_execution_b200 import _lib, configs, programs
from test_front_end import marginals, rule_weights


def random_tree(size_a, seed, depth=5, reach=2):
  rng = numpy.random.default_rng(seed)
  nodes = []  # (kind, a, b, c, children, weights)

  def grow(level):
    me = len(nodes)
    nodes.append(None)
    roll = rng.random()
    if level >= depth or roll < 0.12:
      nodes[me] = (programs.END, 0, 0, 0, [], [])
    elif roll < 0.55:
      kids = [grow(level + 1) for _ in range(size_a)]
      nodes[me] = (programs.READ, int(rng.integers(0, 2)), int(rng.integers(-reach, reach + 1)), 0, kids, [])
    elif roll < 0.8:
      kids = [grow(level + 1)]
      nodes[me] = (programs.WRITE, int(rng.integers(0, 2)), int(rng.integers(-reach, reach + 1)),
                   int(rng.integers(0, size_a)), kids, [])
    else:
      n_opt = int(rng.integers(2, 4))
      weights = rng.random(n_opt)
      if rng.random() < 0.25:
        weights[int(rng.integers(0, n_opt))] = 0.0  # an option that is never taken but never pruned
      kids = [grow(level + 1) for _ in range(n_opt)]
      nodes[me] = (programs.PICK, n_opt, 0, 0, kids, weights.tolist())
    return me
  grow(0)
  n = len(nodes)
  tree = dict(kind=numpy.zeros(n, numpy.int32), a=numpy.zeros(n, numpy.int32), b=numpy.zeros(n, numpy.int32),
              c=numpy.zeros(n, numpy.int32), first_child=numpy.zeros(n, numpy.int32),
              first_weight=numpy.zeros(n, numpy.int32))
  child, weight = [], []
  for i, (kind, a, b, c, kids, wts) in enumerate(nodes):
    tree['kind'][i], tree['a'][i], tree['b'][i], tree['c'][i] = kind, a, b, c
    tree['first_child'][i], tree['first_weight'][i] = len(child), len(weight)
    child.extend(kids)
    weight.extend(wts)
  tree['child'] = numpy.array(child, dtype=numpy.int32)
  tree['weight'] = numpy.array(weight, dtype=numpy.float64)
  return tree


CASES = [(2, 101, 6), (2, 102, 5), (3, 103, 5), (3, 104, 4), (4, 105, 4), (5, 106, 3), (3, 107, 5), (2, 108, 6)]


def tables(size_a, cl_k, seed):
  yield configs.dirichlet_product_table(size_a, cl_k, seed)
  p = configs.markov_table(size_a, cl_k, seed + 1)
  yield p
  rng = numpy.random.default_rng(seed)
  q = p.copy()
  q[rng.random(q.size) < 0.3] = 0.0  # pruned branches
  yield q / q.sum()


@pytest.mark.parametrize('size_a,seed,depth', CASES)
def test_front_end_matches_oracle_on_random_programs(oracle, size_a, seed, depth):
  tree = random_tree(size_a, seed, depth)
  tag = f'rnd-{seed}'
  _lib.register_program(tag, size_a, tree)
  oracle.register_program(tag, size_a, tree)
  for cl_k in (1, 2, 3, 5):
    table = _lib.rule_table(tag, cl_k)
    for p in tables(size_a, cl_k, seed):
      w = rule_weights(table, marginals(p, size_a, cl_k))
      prob, info = oracle.worlds(tag, cl_k, p)
      changed = (info[:, 1] != info[:, 2]) | (info[:, 4] != info[:, 5])
      # the oracle prunes zero-probability worlds, the table keeps them with weight 0
      assert sorted(x for x in w.tolist() if x > 0) == sorted(x for x in prob[changed].tolist() if x > 0)
    lit = oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.LITERAL)
    mer = oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED)
    assert abs(lit - mer).max() <= 1e-13 * max(abs(lit).max(), 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize('size_a,seed,depth', CASES)
def test_gpu_matches_oracle_on_random_programs(oracle, size_a, seed, depth):
  from chemical_kinetics_and_program_execution_b200 import markov_tapes as mt
  from test_gpu_parity import gross_flux
  tree = random_tree(size_a, seed, depth)
  tag = f'rnd-gpu-{seed}'
  _lib.register_program(tag, size_a, tree)
  oracle.register_program(tag, size_a, tree)
  for cl_k in (1, 2, 4, 6 if size_a <= 3 else 5):
    f = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)
    for p in tables(size_a, cl_k, seed):
      got = f(p, 0.0)
      want = oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED)
      # per state, against the gross flux through it; these programs put up to a few hundred terms
      # on a state, and the oracle's own two summation orders already differ by 5e-15 of the largest
      # entry, so the bound is 1e-13 instead of the 1e-14 used for the shipped problems
      assert (abs(got - want) <= 1e-13 * gross_flux(oracle, tag, cl_k, p) + 1e-300).all()
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)


@pytest.mark.parametrize('size_a,seed,depth', CASES)
def test_retraced_tree_is_the_same_program(oracle, size_a, seed, depth):
  """The C++ tracer (which feeds the Monte-Carlo simulator) run on a random tree's body gives a
  tree without the reads of already known cells; both must be the same program to the oracle, bit for
  bit, and tracing the traced tree again must reproduce it."""
  tree = random_tree(size_a, seed, depth)
  tag = f'rnd-retrace-{seed}'
  _lib.register_program(tag, size_a, tree)
  traced = _lib.program_tree(tag)
  assert traced['kind'].size <= tree['kind'].size + 1
  oracle.register_program(tag, size_a, tree)
  oracle.register_program(tag + '-t', size_a, traced)
  _lib.register_program(tag + '-t', size_a, traced)
  again = _lib.program_tree(tag + '-t')
  for key in traced:
    assert numpy.array_equal(traced[key], again[key]), key
  for cl_k in (1, 3, 4):
    for p in tables(size_a, cl_k, seed):
      for mode in (oracle.LITERAL, oracle.MERGED):
        assert numpy.array_equal(oracle.compute_dy_dt(tag, cl_k, p, mode=mode),
                                 oracle.compute_dy_dt(tag + '-t', cl_k, p, mode=mode))
    a, b = _lib.rule_table(tag, cl_k), _lib.rule_table(tag + '-t', cl_k)
    for key in ('rule_ptr', 'step_kind', 'step_len', 'step_long', 'step_short', 'step_prob', 'seed_len', 'seed_orig', 'seed_adj'):
      assert numpy.array_equal(a[key], b[key]), key
