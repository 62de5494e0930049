"""CPU tests of programs given as data (programs.py, tapes_register_program): a Python restatement
of a reference body must produce the flux-rule table of the compiled body, and the oracle must agree
on a program that exists in neither."""

import math

import numpy
import pytest

from chemical_kinetics_and_program_execution_b200 import _lib, configs, programs
from test_front_end import marginals, rule_weights

P, D = False, True


def decay(tape):  # framework/problems.scm:22-26
  if tape.get(D, 0) == 1:
    tape.set(D, 0, 0)


def ferromagnet(tape):  # framework/problems.scm:30-55
  J, h, beta = 1.0, -0.25, 1.0
  mid, lft, rgt = tape.get(D, 0), tape.get(D, -1), tape.get(D, +1)
  bonds = (1 if lft == mid else -1) + (1 if mid == rgt else -1)
  coupling = math.exp(-(beta * J * (4 + 2 * bonds)))
  aligned = (h > 0) == (mid == 1)
  field = math.exp(-(2 * beta * abs(h))) if aligned else 1.0
  flip = coupling * field
  if tape.choose([flip, 1 - flip]) == 0:
    tape.set(D, 0, 0 if mid == 1 else 1)


def copolymerization(tape):  # framework/problems.scm:63-85
  O, A, M, N = 0, 1, 2, 3
  p0 = tape.get(P, 0)
  if p0 == O:
    return
  if not (tape.get(P, -1) == O and tape.get(P, +1) == O):
    return
  under = tape.get(D, 0)
  amine = lambda s: s in (M, N)
  if not ((p0 == A and amine(under)) or (under == A and amine(p0))):
    return
  side = tape.choose_value([(1.0, -1), (1.0, +1)])
  if tape.get(D, side) != O or tape.get(D, 2 * side) != O:
    return
  tape.set(P, 0, O)
  tape.set(D, side, p0)


def relay(tape):
  """Not in the reference: a program-tape token 2 copies the data cell under the head one step to
  the right with a symbol-dependent rate and is used up; token 1 erases the cell to its left."""
  token = tape.get(P, 0)
  if token == 2:
    here = tape.get(D, 0)
    if tape.get(D, 1) != here and tape.choose([0.2 + 0.1 * here, 0.5, 0.3 - 0.1 * here]) == 0:
      tape.set(D, 1, here)
      tape.set(P, 0, 0)
  elif token == 1 and tape.get(D, -1) != 0:
    if tape.choose_value([(0.25, True), (0.75, False)]):
      tape.set(D, -1, 0)


def same_table(x, y):
  for key in ('rule_ptr', 'step_kind', 'step_len', 'step_long', 'step_short', 'seed_len', 'seed_orig', 'seed_adj'):
    assert numpy.array_equal(x[key], y[key]), key
  assert x['leaf_worlds'] == y['leaf_worlds']
  assert numpy.array_equal(x['step_prob'], y['step_prob'])  # the same libm exp on both sides


@pytest.mark.parametrize('body,tag,size_a,ks', [(decay, 'ex1-radioactive-decay', 2, (1, 3, 5)),
                                                (ferromagnet, 'ex2-ferromagnetic-chain', 2, (2, 3, 7)),
                                                (copolymerization, 'ex3-copolymerization', 4, (2, 4, 6))])
def test_python_restatement_equals_compiled_body(oracle, body, tag, size_a, ks):
  tree = programs.trace(body, size_a)
  _lib.register_program('py-' + tag, size_a, tree)
  oracle.register_program('py-' + tag, size_a, tree)
  for cl_k in ks:
    same_table(_lib.rule_table('py-' + tag, cl_k), _lib.rule_table(tag, cl_k))
    p = configs.markov_table(size_a, cl_k, 3)
    for mode in (oracle.LITERAL, oracle.MERGED):
      assert numpy.array_equal(oracle.compute_dy_dt('py-' + tag, cl_k, p, mode=mode),
                               oracle.compute_dy_dt(tag, cl_k, p, mode=mode))


def test_new_program_matches_oracle_worlds(oracle):
  tree = programs.trace(relay, 3)
  assert tree['kind'][-1] == programs.END and (tree['kind'][:-1] != programs.END).all()
  _lib.register_program('py-relay', 3, tree)
  oracle.register_program('py-relay', 3, tree)
  for cl_k in (1, 2, 4):
    table = _lib.rule_table('py-relay', cl_k)
    p = configs.dirichlet_product_table(3, cl_k, 5)
    w = rule_weights(table, marginals(p, 3, cl_k))
    prob, info = oracle.worlds('py-relay', cl_k, p)
    changed = (info[:, 1] != info[:, 2]) | (info[:, 4] != info[:, 5])
    assert table['leaf_worlds'] == len(prob)
    assert sorted(w.tolist()) == sorted(prob[changed].tolist())
    dy = oracle.compute_dy_dt('py-relay', cl_k, p, mode=oracle.MERGED)
    assert abs(dy.sum()) <= 1e-15 and abs(dy).max() > 0


def test_malformed_programs_are_rejected():
  calls = [0]

  def moody(tape):  # behaves differently from run to run
    calls[0] += 1
    if calls[0] % 2:
      tape.get(D, 0)
      tape.get(D, 1)
    else:
      tape.choose([1, 1])
  with pytest.raises(ValueError, match='deterministic'):
    programs.trace(moody, 2)
  with pytest.raises(ValueError, match='alphabet'):
    programs.trace(lambda tape: tape.set(D, 0, 7), 3)
  tree = programs.trace(decay, 2)
  bad = dict(tree, child=tree['child'].copy())
  bad['child'][0] = 0  # a cycle
  with pytest.raises(RuntimeError, match='after its parent'):
    _lib.register_program('py-bad', 2, bad)
  far = programs.trace(lambda tape: tape.get(D, 400), 2)
  with pytest.raises(RuntimeError, match='too far'):
    _lib.register_program('py-far', 2, far)
  assert _lib.load().tapes_alphabet_size(b'py-bad') == -1
