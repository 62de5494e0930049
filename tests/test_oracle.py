"""CPU tests of the oracle against the reference's known answers and its own invariants."""

import os

import numpy
import pytest
import scipy.integrate

from conftest import dense
from chemical_kinetics_and_program_execution_b200 import configs

TAGS = [('ex1-radioactive-decay', 2, 4), ('ex2-ferromagnetic-chain', 2, 5),
        ('ex3-copolymerization', 4, 4), ('ex3var1-copolymerization', 4, 4),
        ('ex3var2-copolymerization', 4, 4), ('ex4-chemical-turing', 9, 3),
        ('ex4var1-chemical-turing', 9, 3), ('ex4var2-chemical-turing', 10, 3),
        ('ex5-msrtf-machine', 5, 4), ('ex5var1-msrtf-machine', 5, 4)]


def test_canary_exact(oracle, known_answers):
  # framework/markov_tapes.py:357-365: exact equality
  for mode in (oracle.LITERAL, oracle.MERGED):
    out = oracle.compute_dy_dt(known_answers['canary_tag'], known_answers['canary_cl_k'],
                               known_answers['canary_p'], mode=mode)
    assert out.tolist() == known_answers['canary_dy_dt']


def test_canary_call_counts(oracle, known_answers):
  # 16 literal accumulate calls for 12 distinct (src, dst) pairs
  _, c = oracle.compute_dy_dt(known_answers['canary_tag'], 3, known_answers['canary_p'],
                              mode=oracle.LITERAL, want_counters=True)
  assert c['acc_calls'] == 16
  src, dst, w = oracle.terms(known_answers['canary_tag'], 3, known_answers['canary_p'],
                             mode=oracle.MERGED)
  assert len(src) == 12 and len(set(zip(src.tolist(), dst.tolist()))) == 12


@pytest.mark.parametrize('tag,size_a,cl_k', TAGS)
def test_literal_equals_merged_and_conserves(oracle, tag, size_a, cl_k):
  for seed, make in ((1, configs.dirichlet_product_table), (2, configs.markov_table)):
    p = make(size_a, cl_k, seed)
    lit = oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.LITERAL)
    mer = oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED)
    scale = abs(lit).max() + 1e-300
    assert abs(lit - mer).max() <= 1e-13 * scale
    assert abs(lit.sum()) <= 1e-13 * abs(lit).sum() + 1e-300  # every term adds +w and -w


def test_unknown_tag_is_an_error(oracle):
  with pytest.raises(KeyError):
    oracle.compute_dy_dt('no-such-problem', 3, numpy.ones(8) / 8)


def test_p0_builders_match_reference_generators(p0_fixtures):
  fx = p0_fixtures
  for k in range(3, 8):
    want = dense(fx[f'ex2_k{k}_idx'], fx[f'ex2_k{k}_val'], 2 ** k)
    assert numpy.array_equal(configs.ex2_p0(k, p_pair=1 / 250), want)
  assert numpy.array_equal(configs.ex3_p0(), dense(fx['ex3_k6_idx'], fx['ex3_k6_val'], 4 ** 6))
  assert numpy.array_equal(configs.ex4_p0(powered_fraction=0.04),
                           dense(fx['ex4_a_idx'], fx['ex4_a_val'], 9 ** 5))
  assert numpy.array_equal(configs.ex4_p0(powered_fraction=0.01),
                           dense(fx['ex4_b_idx'], fx['ex4_b_val'], 9 ** 5))
  assert numpy.array_equal(configs.ex4var2_p0(), dense(fx['ex4var2_idx'], fx['ex4var2_val'], 10 ** 5))
  assert numpy.array_equal(configs.ex5_p0(), dense(fx['ex5_idx'], fx['ex5_val'], 5 ** 5))


def test_shipped_p0_are_shift_consistent(p0_fixtures):
  fx = p0_fixtures
  for name, size_a, k in (('ex3_k6', 4, 6), ('ex4_a', 9, 5), ('ex5', 5, 5), ('ex2_k7', 2, 7)):
    p = dense(fx[name + '_idx'], fx[name + '_val'], size_a ** k).reshape([size_a] * k)
    assert abs(p.sum(axis=0) - p.sum(axis=-1)).max() < 1e-15


def test_ex4_end_points_match_reference(oracle, known_answers, p0_fixtures):
  # examples/ex4_chemical_turing.py:150-170, the reference's only recorded trajectory values
  p0 = dense(p0_fixtures['ex4_b_idx'], p0_fixtures['ex4_b_val'], 9 ** 5)
  f = oracle.get_dy_dt(tag='ex4-chemical-turing', size_a=9, cl_k=5, mode=oracle.MERGED)
  sol = scipy.integrate.solve_ivp(lambda t, y: f(y, t), (0.0, 2000.0), p0, t_eval=[0.0, 2000.0],
                                  rtol=1e-13, atol=1e-13, method='DOP853')
  spd = sol.y[:, -1].reshape([9] * 5)
  got = [float(spd[(Ellipsis,) + tuple(s)].sum()) for s in known_answers['ex4_observables']]
  for g, w in zip(got, known_answers['ex4_p0_b_t2000']):
    assert abs(g - w) <= 1e-12 * abs(w)


def test_rule_table_problem(oracle):
  rules = configs.random_rule_set(4, 6, seed=3)
  oracle.register_rules('rt-test', 4, rules)
  p = configs.dirichlet_product_table(4, 4, 5)
  lit = oracle.compute_dy_dt('rt-test', 4, p, mode=oracle.LITERAL)
  mer = oracle.compute_dy_dt('rt-test', 4, p, mode=oracle.MERGED)
  assert abs(lit).max() > 0
  assert abs(lit - mer).max() <= 1e-13 * abs(lit).max()
  assert abs(lit.sum()) <= 1e-13 * abs(lit).sum()


def test_autocatalysis_rule_set(oracle):
  rules = configs.autocatalysis_rule_set()
  oracle.register_rules('autocat-cpu', 4, rules)
  p = configs.markov_table(4, 5, 7)
  lit = oracle.compute_dy_dt('autocat-cpu', 5, p, mode=oracle.LITERAL)
  mer = oracle.compute_dy_dt('autocat-cpu', 5, p, mode=oracle.MERGED)
  assert abs(lit).max() > 0
  assert abs(lit - mer).max() <= 1e-13 * abs(lit).max()
  assert abs(lit.sum()) <= 1e-13 * abs(lit).sum()


def test_ex3_long_chain_golden_is_what_the_oracle_gives(oracle):
  """tests/golden/ex3_k12_trajectory.npz (BASELINE config 3, used by the GPU suite) is the merged-mode
  oracle through SciPy's DOP853; regenerating its first output times reproduces it bit for bit."""
  import scipy.integrate
  from conftest import GOLDEN
  from make_golden_round2 import EX3_SEQS, seq_sum
  from chemical_kinetics_and_program_execution_b200 import configs
  gold = numpy.load(os.path.join(GOLDEN, 'ex3_k12_trajectory.npz'))
  f = oracle.get_dy_dt(tag='ex3-copolymerization', size_a=4, cl_k=12, mode=oracle.MERGED)
  # t_eval does not influence the steps, and the steps up to t = 2 do not depend on what follows:
  # a terminal event ends the run once t = 2 has been passed
  def passed(t, y):
    return t - 2.5
  passed.terminal = True
  sol = scipy.integrate.solve_ivp(lambda t, y: f(y, t), (0.0, 10.0), configs.ex3_p0(12), t_eval=[0.0, 1.0, 2.0],
                                  method='DOP853', rtol=1e-10, atol=1e-10, events=passed)
  for i in range(3):
    got = [seq_sum(sol.y[:, i], 4, 12, s) for s in EX3_SEQS]
    assert got == gold['observables'][i].tolist()
