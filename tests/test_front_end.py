"""CPU tests of the product's host-side rule front end (csrc/rules.cpp, csrc/problems.cpp):
the flux-rule table must describe exactly the leaf worlds the oracle visits."""

import numpy
import pytest

from chemical_kinetics_and_program_execution_b200 import _lib, configs
from test_oracle import TAGS


def marginals(p, size_a, cl_k):
  tabs = {cl_k: numpy.asarray(p, dtype=numpy.float64).ravel()}
  for L in range(cl_k - 1, -1, -1):
    src = tabs[L + 1].reshape(-1, size_a)
    acc = numpy.zeros(src.shape[0])
    for j in range(size_a):  # sequential, like tm.scm:380-384
      acc = acc + src[:, j]
    tabs[L] = acc
  return tabs


def rule_weights(table, tabs):
  out = []
  for r in range(len(table['rule_ptr']) - 1):
    w = 1.0
    for s in range(table['rule_ptr'][r], table['rule_ptr'][r + 1]):
      if table['step_kind'][s] == 0:
        L = table['step_len'][s]
        p_here = max(0.0, tabs[L][table['step_long'][s]])
        p_marg = tabs[L - 1][table['step_short'][s]]
        rel = 0.0 if p_here == 0 else p_here / max(p_here, p_marg)
        w = w * rel
        if not w > 0:
          w = 0.0
          break
      else:
        w = max(0.0, table['step_prob'][s]) * w
    out.append(w)
  return numpy.array(out)


@pytest.mark.parametrize('tag,size_a,cl_k', TAGS + [('ex4-chemical-turing', 9, 2),
                                                   ('ex2-ferromagnetic-chain', 2, 2),
                                                   ('ex5-msrtf-machine', 5, 3)])
def test_rule_table_matches_oracle_worlds(oracle, tag, size_a, cl_k):
  assert _lib.load().tapes_alphabet_size(tag.encode()) == size_a
  table = _lib.rule_table(tag, cl_k)
  p = configs.dirichlet_product_table(size_a, cl_k, 11)
  tabs = marginals(p, size_a, cl_k)
  w = rule_weights(table, tabs)
  prob, info = oracle.worlds(tag, cl_k, p)
  changed = (info[:, 1] != info[:, 2]) | (info[:, 4] != info[:, 5])
  assert table['leaf_worlds'] == len(prob)  # full support: nothing is pruned
  want = sorted((tuple(info[i].tolist()), prob[i]) for i in numpy.nonzero(changed)[0])
  got = sorted(((int(table['seed_len'][r, 0]), int(table['seed_orig'][r, 0]), int(table['seed_adj'][r, 0]),
                 int(table['seed_len'][r, 1]), int(table['seed_orig'][r, 1]), int(table['seed_adj'][r, 1])),
                w[r]) for r in range(len(w)))
  assert [g[0] for g in got] == [x[0] for x in want]
  for g, x in zip(got, want):
    assert g[1] == x[1]  # same operations in the same order: bit-identical


def test_rule_set_problem_matches_oracle(oracle):
  rules = configs.random_rule_set(5, 7, seed=4)
  oracle.register_rules('rt-front', 5, rules)
  _lib.register_rules('rt-front', 5, rules)
  table = _lib.rule_table('rt-front', 4)
  p = configs.markov_table(5, 4, 2)
  w = rule_weights(table, marginals(p, 5, 4))
  prob, info = oracle.worlds('rt-front', 4, p)
  changed = (info[:, 1] != info[:, 2]) | (info[:, 4] != info[:, 5])
  assert sorted(w.tolist()) == sorted(prob[changed].tolist())


@pytest.mark.parametrize('tag,size_a,cl_k', TAGS)
def test_rule_parts_deal_every_rule_once_and_balanced(oracle, tag, size_a, cl_k):
  """The dealing behind tapes_model_part: the estimated cost of a flux rule is its number of
  distinct flux terms at full support (checked against the oracle's merged term list), every rule
  has exactly one owner, and the longest-processing-time bound holds."""
  p = configs.dirichlet_product_table(size_a, cl_k, 11)
  n_terms = len(oracle.terms(tag, cl_k, p, mode=oracle.MERGED)[0])
  n_rules = len(_lib.rule_table(tag, cl_k)['rule_ptr']) - 1
  for n_parts in (1, 2, 3, 8):
    owner, cost = _lib.rule_parts(tag, cl_k, n_parts)
    assert len(owner) == n_rules and cost.sum() == n_terms and (cost > 0).all()
    assert owner.min() >= 0 and owner.max() < n_parts
    load = numpy.bincount(owner, weights=cost, minlength=n_parts)
    assert load.max() <= cost.sum() / n_parts + cost.max()
    again, _ = _lib.rule_parts(tag, cl_k, n_parts)
    assert numpy.array_equal(owner, again)  # every rank computes the same dealing
  assert (_lib.rule_parts(tag, cl_k, 1)[0] == 0).all()


def test_rule_parts_errors():
  with pytest.raises(RuntimeError):
    _lib.rule_parts('nope', 3, 2)
  with pytest.raises(RuntimeError):
    _lib.rule_parts('ex2-ferromagnetic-chain', 3, 0)


def test_unknown_tag():
  assert _lib.load().tapes_alphabet_size(b'nope') == -1
  with pytest.raises(RuntimeError):
    _lib.rule_table('nope', 3)
