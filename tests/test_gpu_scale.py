"""GPU parity at BASELINE.json's larger configurations: oracle comparisons where the CPU port
finishes in seconds, size-independent properties (conservation, additivity over rule shards) at
the full 10^8-state size."""

import numpy
import pytest

pytestmark = pytest.mark.gpu

from chemical_kinetics_and_program_execution_b200 import configs, parallel  # noqa: E402


@pytest.fixture(scope='module')
def mt():
  from chemical_kinetics_and_program_execution_b200 import markov_tapes
  return markov_tapes


@pytest.fixture(scope='module')
def device():
  from chemical_kinetics_and_program_execution_b200 import device as dev
  return dev


def close(got, want, tol=1e-12):
  scale = abs(want).max()
  assert abs(got - want).max() <= tol * scale, (abs(got - want).max(), scale)


def test_ex3_long_chain(mt, oracle):
  """BASELINE config 3: ex3 with cl_k = 12 (1.68e7 states)."""
  f = mt.get_dy_dt(tag='ex3-copolymerization', size_a=4, cl_k=12)
  p0 = configs.ex3_p0(12)  # the shipped generator at k = 12: sparse, heavily pruned
  close(f(p0, 0.0), oracle.compute_dy_dt('ex3-copolymerization', 12, p0, mode=oracle.LITERAL))
  p = configs.dirichlet_product_table(4, 12, 3)  # full support: every term is live
  got = f(p, 0.0)
  close(got, oracle.compute_dy_dt('ex3-copolymerization', 12, p, mode=oracle.MERGED))
  assert abs(got.sum()) <= 1e-13 * abs(got).sum()


def test_ex3_long_chain_as_composite_model(mt, oracle, monkeypatch):
  """The same configuration split into several structures (as a forest above 2^31 nodes would be):
  exercises the host-buffer path that returns the result in row blocks while later blocks are
  still being summed over the structures."""
  p = configs.dirichlet_product_table(4, 12, 3)
  mt.u_lib.tapes_release_model(b'ex3-copolymerization', 12)
  monkeypatch.setenv('TAPES_MAX_PART_TERMS', '30000000')
  try:
    stats = mt.model_stats(tag='ex3-copolymerization', cl_k=12)
    assert stats['structures'] >= 3
    got = mt.get_dy_dt(tag='ex3-copolymerization', size_a=4, cl_k=12)(p, 0.0)
    close(got, oracle.compute_dy_dt('ex3-copolymerization', 12, p, mode=oracle.MERGED))
  finally:
    mt.u_lib.tapes_release_model(b'ex3-copolymerization', 12)


def test_autocatalysis_tape(mt, oracle):
  """BASELINE config 2: our tape restatement of the autocatalysis chemistry at ~10^6 states."""
  rules = configs.autocatalysis_rule_set()
  oracle.register_rules('autocatalysis-tape', 4, rules)
  mt.register_rule_set('autocatalysis-tape', 4, rules)
  f = mt.get_dy_dt(tag='autocatalysis-tape', size_a=4, cl_k=10)
  for p in (configs.product_table([0.5, 0.3, 0.1, 0.1], 10), configs.markov_table(4, 10, 4)):
    got = f(p, 0.0)
    close(got, oracle.compute_dy_dt('autocatalysis-tape', 10, p, mode=oracle.MERGED))
    assert abs(got.sum()) <= 1e-13 * abs(got).sum()
  # a short HBM-resident integration keeps the table a probability distribution
  p0 = configs.product_table([0.5, 0.3, 0.1, 0.1], 10)
  seqs = [[1], [2], [3], [2, 2], [3, 3]]
  series = mt.ode_integrate_device(tag='autocatalysis-tape', size_a=4, cl_k=10, p0=p0, ts=numpy.linspace(0, 5, 6),
                                   rtol=1e-8, atol=1e-10, observables=seqs + [[0]], return_states=False)
  assert (series > -1e-12).all()
  assert abs(series[:, [0, 1, 2, 5]].sum(axis=1) - 1).max() < 1e-9  # single-symbol marginals sum to 1
  assert series[-1, 3] > series[0, 3]  # autocatalysed A-dimers grow from the seeded A halves


def test_full_size_synthetic_properties(mt, device, oracle):
  """BASELINE config 5 at its full size (A = 10, k = 8, 10^8 states): conservation, additivity
  over rule shards, and one rule against the CPU port."""
  import torch
  import bench
  size_a, cl_k, n_rules = 10, 8, 3
  rules = configs.random_rule_set(size_a, n_rules, seed=11)
  mt.register_rule_set('scale-full', size_a, rules)
  full = device.DeviceModel('scale-full', cl_k)
  p = bench.device_product_table(size_a, cl_k, 5, torch.device('cuda'))
  want = full.rhs(p).clone()
  torch.cuda.synchronize()
  assert float(want.sum().abs()) <= 1e-12 * float(want.abs().sum())  # every term adds +w and -w
  total = torch.zeros_like(want)
  for r in range(n_rules):
    mt.register_rule_set(f'scale-part{r}', size_a, parallel.split_rule_set(rules, n_rules, r))
    part = device.DeviceModel(f'scale-part{r}', cl_k)
    total += part.rhs(p)
    if r == 0:
      oracle.register_rules('scale-part0', size_a, parallel.split_rule_set(rules, n_rules, 0))
      cpu = oracle.compute_dy_dt('scale-part0', cl_k, p.cpu().numpy(), mode=oracle.MERGED)
      torch.cuda.synchronize()
      close(part.rhs(p).cpu().numpy(), cpu)
    mt.u_lib.tapes_release_model(f'scale-part{r}'.encode(), cl_k)
  torch.cuda.synchronize()
  assert float((total - want).abs().max()) <= 1e-13 * float(want.abs().max())
  mt.u_lib.tapes_release_model(b'scale-full', cl_k)


def test_ex3_long_chain_trajectory(mt):
  """BASELINE config 3 end to end: examples/ex3_copolymerization.py:38-64
  with cl_k = 12 (1.68e7 states) integrated to a fixed horizon, t in [0, 10], DOP853 at
  rtol = atol = 1e-10, table resident in HBM, the observables of ex3_copolymerization.py:112-118 read
  on the device.  Golden: the CPU oracle through SciPy's DOP853 with the same settings
  (tests/golden/make_golden_round2.py).  Same steps on both sides (equal number of right-hand
  sides), so the trajectories differ through dy/dt rounding only: 1e-12 relative."""
  import os
  import time
  from conftest import GOLDEN
  from make_golden_round2 import EX3_SEQS
  gold = numpy.load(os.path.join(GOLDEN, 'ex3_k12_trajectory.npz'))
  p0 = configs.ex3_p0(12)
  mt.model_stats(tag='ex3-copolymerization', cl_k=12)  # build outside the timed part
  t0 = time.perf_counter()
  series, stats = mt.ode_integrate_device(tag='ex3-copolymerization', size_a=4, cl_k=12, p0=p0, ts=gold['ts'],
                                          rtol=1e-10, atol=1e-10, observables=EX3_SEQS, return_states=False,
                                          want_stats=True)
  seconds = time.perf_counter() - t0
  print(f'ex3 cl_k=12 to t=10: {seconds:.3f} s, {stats["nfev"]} right-hand sides, {stats["accepted"]} steps')
  # solve_ivp counts the right-hand sides of the steps only; the dense output adds 3 per step that has
  # an output time in it on both sides
  assert stats['nfev'] == int(gold['nfev'][0]), (stats, gold['nfev'])
  want = gold['observables']
  assert series.shape == want.shape
  assert (abs(series - want) <= 1e-12 * abs(want) + 1e-300).all(), abs(series / numpy.where(want == 0, 1, want) - 1).max()
  states = mt.ode_integrate_device(tag='ex3-copolymerization', size_a=4, cl_k=12, p0=p0, ts=[0.0, 10.0],
                                   rtol=1e-10, atol=1e-10)
  end = numpy.zeros(4 ** 12)
  end[gold['end_idx']] = gold['end_val']
  close(states[-1], end)
  assert (abs(states[-1] - end) <= 1e-12 * abs(end) + 1e-18).all()
  assert abs(states[-1].sum() - 1) < 1e-12


def test_pinned_result_pool_and_pageable_buffers(mt):
  """Large tables through the host-buffer entry point: the mirror's get_dy_dt returns arrays over
  page-locked memory from a small pool (reused once the caller drops them, ordinary arrays when
  all are still held); the reference's own binding hands over pageable NumPy buffers, which go
  through the threaded staging (csrc/hostcopy.h).  Same bits every way."""
  from test_abi import reference_binding, reference_dy_dt
  f = mt.get_dy_dt(tag='ex3-copolymerization', size_a=4, cl_k=12)
  p = configs.dirichlet_product_table(4, 12, 3)
  first = f(p, 0.0)
  address = first.ctypes.data
  keep = first.copy()
  del first
  again = f(p, 0.0)
  assert again.ctypes.data == address and numpy.array_equal(again, keep)  # the buffer came back
  held = [again] + [f(p, 0.0) for _ in range(6)]  # more than the pool holds
  assert all(numpy.array_equal(x, keep) for x in held)
  assert len({x.ctypes.data for x in held}) == len(held)
  out = reference_dy_dt(reference_binding(), 'ex3-copolymerization', 12, p)  # pageable in, pageable out
  assert numpy.array_equal(out, keep)
  assert abs(keep.sum()) <= 1e-13 * abs(keep).sum()
