"""GPU parity tests: the CUDA path (called through the C ABI) against the CPU oracle.

Tolerances (stated per BASELINE.json's north_star): the flux-term set is compared exactly
(integer indices); dy/dt is compared per state to 1e-14 of the gross flux through that state
(sums of ~1e1..1e2 products of ~10 ratios each, evaluated in a different order than the CPU's
depth-first walk); trajectories
driven by the same SciPy stepper are compared to 1e-12 relative.
"""

import json
import os

import numpy
import pytest
import scipy.integrate

from conftest import dense

pytestmark = pytest.mark.gpu

from chemical_kinetics_and_program_execution_b200 import configs  # noqa: E402
from test_oracle import TAGS  # noqa: E402



@pytest.fixture(scope='module')
def mt():
  from chemical_kinetics_and_program_execution_b200 import markov_tapes
  return markov_tapes


@pytest.fixture(scope='module')
def device():
  from chemical_kinetics_and_program_execution_b200 import device as dev
  return dev


def test_canary_exact(mt, known_answers):
  # the reference's load-time known-answer test, framework/markov_tapes.py:357-365
  f = mt.get_dy_dt(tag=known_answers['canary_tag'], size_a=2, cl_k=3)
  assert f(numpy.array(known_answers['canary_p']), 0.0).tolist() == known_answers['canary_dy_dt']


def gross_flux(oracle, tag, cl_k, p):
  """Per state, the sum of |w| over the flux terms touching it (the scale rounding errors live on:
  near a steady state dy/dt is a small difference of large in- and out-flows)."""
  src, dst, w = oracle.terms(tag, cl_k, p, mode=oracle.MERGED)
  g = numpy.zeros(numpy.asarray(p).size)
  numpy.add.at(g, src, abs(w))
  numpy.add.at(g, dst, abs(w))
  return g


def assert_rhs_close(got, want, gross, tolerance=1e-14):
  # dy/dt within ~1e-15 * sum|terms|; 1e-14 allows ~45 ulp: each term is a product of
  # up to ~25 ratios (ex5 reads 4 + 3 cells, then k window extensions), summed in another order
  err = abs(got - want)
  assert (err <= tolerance * gross + 1e-300).all(), (err.max(), gross.max(), abs(want).max())


def check_rhs(f, oracle, tag, cl_k, p, mode):
  assert_rhs_close(f(p, 0.0), oracle.compute_dy_dt(tag, cl_k, p, mode=mode),
                   gross_flux(oracle, tag, cl_k, p))


@pytest.mark.parametrize('tag,size_a,cl_k', TAGS + [('ex2-ferromagnetic-chain', 2, 2),
                                                   ('ex2-ferromagnetic-chain', 2, 3),
                                                   ('ex4-chemical-turing', 9, 2),
                                                   ('ex1-radioactive-decay', 2, 1),
                                                   ('ex3-copolymerization', 4, 6),
                                                   ('ex5-msrtf-machine', 5, 5)])
def test_rhs_matches_oracle(mt, oracle, tag, size_a, cl_k):
  f = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)
  for seed, make in ((1, configs.dirichlet_product_table), (2, configs.markov_table)):
    p = make(size_a, cl_k, seed)
    check_rhs(f, oracle, tag, cl_k, p, oracle.MERGED)


@pytest.mark.parametrize('seed', range(48))
def test_random_rule_sets_and_sparse_tables_match_the_oracle(mt, oracle, seed):
  """Differential test over shapes the fixed cases do not reach: alphabet 2..6, window 2..7, 1..6
  random rewrite rules (spans 1..3, with and without a catalyst on the other tape), and tables
  with a third to a half of their entries exactly zero (the pruning branches, tm.scm:1316, 1350,
  1373) next to full-support ones.  dy/dt through the reference's host entry point against the
  oracle in both modes, and the sum over all states must vanish."""
  rng = numpy.random.default_rng(1000 + seed)
  size_a, cl_k, n_rules = int(rng.integers(2, 7)), int(rng.integers(2, 8)), int(rng.integers(1, 7))
  while size_a ** cl_k > 50000:
    cl_k -= 1
  rules = configs.random_rule_set(size_a, n_rules, seed=seed, catalyst_fraction=float(rng.random()))
  tag = f'fuzz-{seed}'
  mt.register_rule_set(tag, size_a, rules)
  oracle.register_rules(tag, size_a, rules)
  f = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)
  full = configs.markov_table(size_a, cl_k, seed)
  sparse = full * (rng.random(full.shape) > rng.uniform(0.3, 0.5))
  sparse = sparse / sparse.sum() if sparse.sum() > 0 else full  # a 4-entry table can lose everything
  for p in (full, sparse):
    got = f(p, 0.0)
    gross = gross_flux(oracle, tag, cl_k, p)
    # the GPU adds the parents of a prefix group before multiplying, like the oracle's merged mode; the
    # literal recursion multiplies first, and the oracle's two modes themselves differ by up to 1.03e-14 of
    # the gross flux on these cases (seed 17): 3e-14 against the literal mode
    assert_rhs_close(got, oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED), gross)
    assert_rhs_close(got, oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.LITERAL), gross, tolerance=3e-14)
    assert abs(got.sum()) <= 1e-14 * gross.sum() + 1e-300
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)


def test_rhs_matches_literal_oracle_on_shipped_p0(mt, oracle, p0_fixtures):
  fx = p0_fixtures
  cases = [('ex2-ferromagnetic-chain', 2, k, dense(fx[f'ex2_k{k}_idx'], fx[f'ex2_k{k}_val'], 2 ** k))
           for k in range(3, 8)]
  cases += [('ex3-copolymerization', 4, 6, dense(fx['ex3_k6_idx'], fx['ex3_k6_val'], 4 ** 6)),
            ('ex3var1-copolymerization', 4, 6, dense(fx['ex3_k6_idx'], fx['ex3_k6_val'], 4 ** 6)),
            ('ex3var2-copolymerization', 4, 6, dense(fx['ex3_k6_idx'], fx['ex3_k6_val'], 4 ** 6)),
            ('ex4-chemical-turing', 9, 5, dense(fx['ex4_a_idx'], fx['ex4_a_val'], 9 ** 5)),
            ('ex4var1-chemical-turing', 9, 5, dense(fx['ex4_b_idx'], fx['ex4_b_val'], 9 ** 5)),
            ('ex4var2-chemical-turing', 10, 5, dense(fx['ex4var2_idx'], fx['ex4var2_val'], 10 ** 5)),
            ('ex5-msrtf-machine', 5, 5, dense(fx['ex5_idx'], fx['ex5_val'], 5 ** 5)),
            ('ex5var1-msrtf-machine', 5, 5, dense(fx['ex5_idx'], fx['ex5_val'], 5 ** 5))]
  for tag, size_a, cl_k, p0 in cases:
    f = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)
    check_rhs(f, oracle, tag, cl_k, p0, oracle.LITERAL)


def test_rhs_on_mid_trajectory_states(mt, oracle, trajectories):
  tr = trajectories
  f = mt.get_dy_dt(tag='ex4-chemical-turing', size_a=9, cl_k=5)
  for name in ('a', 'b'):
    for tt in (10, 100, 1000):
      p = dense(tr[f'ex4_{name}_t{tt}_idx'], tr[f'ex4_{name}_t{tt}_val'], 9 ** 5)
      check_rhs(f, oracle, 'ex4-chemical-turing', 5, p, oracle.LITERAL)
  f5 = mt.get_dy_dt(tag='ex5-msrtf-machine', size_a=5, cl_k=5)
  for key in ('ex5_t50', 'ex5_end'):
    check_rhs(f5, oracle, 'ex5-msrtf-machine', 5, tr[key], oracle.MERGED)


def test_rhs_with_slightly_negative_entries(mt, oracle):
  # integrators feed slightly invalid tables (tm.scm:526-539); clamps must behave identically
  rng = numpy.random.default_rng(5)
  p = configs.dirichlet_product_table(4, 5, 3)
  p = p + 1e-9 * rng.standard_normal(p.size) * (rng.random(p.size) < 0.05)
  p[rng.integers(0, p.size, 40)] = 0.0
  p[rng.integers(0, p.size, 10)] = -1e-12
  for tag in ('ex3-copolymerization', 'ex3var2-copolymerization'):
    f = mt.get_dy_dt(tag=tag, size_a=4, cl_k=5)
    check_rhs(f, oracle, tag, 5, p, oracle.MERGED)


@pytest.mark.parametrize('tag,size_a,cl_k', [('__canary_problem_radioactive_decay', 2, 3),
                                            ('ex2-ferromagnetic-chain', 2, 5),
                                            ('ex3var2-copolymerization', 4, 4),
                                            ('ex4-chemical-turing', 9, 3),
                                            ('ex5-msrtf-machine', 5, 4)])
def test_flux_term_set_is_bit_exact(mt, device, oracle, tag, size_a, cl_k):
  """State set / sparsity: the (src, dst) pairs and their weights, canonically sorted."""
  import torch
  model = device.DeviceModel(tag, cl_k)
  p = configs.dirichlet_product_table(size_a, cl_k, 7)  # full support: nothing pruned
  d_p = torch.from_numpy(p).cuda()
  model.rhs(d_p)
  torch.cuda.synchronize()
  src, dst, w = model.terms()
  osrc, odst, ow = oracle.terms(tag, cl_k, p, mode=oracle.MERGED)
  assert model.info['n_terms'] == len(src)

  def canon(s, d, ww):
    order = numpy.lexsort((numpy.arange(len(s)), d, s))
    keys = numpy.stack([s[order], d[order]], axis=1)
    uniq, inv = numpy.unique(keys, axis=0, return_inverse=True)
    tot = numpy.zeros(len(uniq))
    numpy.add.at(tot, inv.ravel(), ww[order])
    return uniq, tot
  gk, gw = canon(src, dst, w)
  ok, ow2 = canon(osrc, odst, ow)
  assert numpy.array_equal(gk, ok)  # bit-exact sparsity
  assert abs(gw - ow2).max() <= 1e-14 * abs(ow2).max()
  # the CSR rows are sorted and the row pointer is monotone
  row_ptr, entries = model.csr()
  assert (numpy.diff(row_ptr) >= 0).all() and row_ptr[-1] == len(entries) == 2 * len(src)
  for r in numpy.nonzero(numpy.diff(row_ptr) > 1)[0][:200]:
    seg = entries[row_ptr[r]:row_ptr[r + 1]]
    assert (numpy.diff(seg.astype(numpy.int64)) > 0).all()


def test_rule_set_problem(mt, oracle):
  rules = configs.random_rule_set(6, 9, seed=8)
  tag = 'rt-gpu'
  oracle.register_rules(tag, 6, rules)
  mt.register_rule_set(tag, 6, rules)
  for cl_k in (2, 4, 5):
    p = configs.markov_table(6, cl_k, 9)
    f = mt.get_dy_dt(tag=tag, size_a=6, cl_k=cl_k)
    check_rhs(f, oracle, tag, cl_k, p, oracle.MERGED)


def test_python_programs(mt, oracle):
  """Problems stated as Python functions (markov_tapes.register_program): a restatement of a
  reference body gives the bits of the compiled body, and a program that exists nowhere else
  matches the oracle interpreting the same tree."""
  import test_programs as tp
  from chemical_kinetics_and_program_execution_b200 import programs
  mt.register_program('py-gpu-ex2', 2, tp.ferromagnet)
  mt.register_program('py-gpu-ex3', 4, tp.copolymerization)
  for tag, ref, size_a, cl_k in (('py-gpu-ex2', 'ex2-ferromagnetic-chain', 2, 7),
                                 ('py-gpu-ex3', 'ex3-copolymerization', 4, 6)):
    p = configs.markov_table(size_a, cl_k, 12)
    got = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)(p, 0.0)
    assert numpy.array_equal(got, mt.get_dy_dt(tag=ref, size_a=size_a, cl_k=cl_k)(p, 0.0))
  mt.register_program('py-gpu-relay', 3, tp.relay)
  oracle.register_program('py-gpu-relay', 3, programs.trace(tp.relay, 3))
  for cl_k in (1, 2, 5, 8):
    for make in (configs.dirichlet_product_table, configs.markov_table):
      p = make(3, cl_k, 14)
      check_rhs(mt.get_dy_dt(tag='py-gpu-relay', size_a=3, cl_k=cl_k), oracle, 'py-gpu-relay', cl_k, p, oracle.MERGED)
  p = configs.markov_table(3, 4, 15)
  check_rhs(mt.get_dy_dt(tag='py-gpu-relay', size_a=3, cl_k=4), oracle, 'py-gpu-relay', 4, p, oracle.LITERAL)
  # re-registering a tag replaces the program and drops the structures built for the old one
  before = mt.get_dy_dt(tag='py-gpu-relay', size_a=3, cl_k=4)(p, 0.0)
  mt.register_program('py-gpu-relay', 3, lambda tape: tape.set(True, 0, 0) if tape.get(True, 0) else None)
  after = mt.get_dy_dt(tag='py-gpu-relay', size_a=3, cl_k=4)(p, 0.0)
  assert not numpy.array_equal(before, after)


def random_program(seed, size_a):
  """A random tape program as a Python body: a decision tree over reads (cells -2..2 of either tape),
  writes and weighted choices, at most four reads deep, that depends on nothing but what it reads and
  chooses (the tree is drawn once, here; the body only walks it)."""
  rng = numpy.random.default_rng(seed)

  def grow(depth, reads):
    kind = rng.choice(['read', 'pick', 'write', 'end'], p=[0.45, 0.15, 0.3, 0.1] if depth < 6 else [0, 0, 0.5, 0.5])
    if kind == 'read' and reads >= 4:
      kind = 'write'
    if kind == 'read':
      return ('read', bool(rng.integers(2)), int(rng.integers(-2, 3)), [grow(depth + 1, reads + 1) for _ in range(size_a)])
    if kind == 'pick':
      ways = int(rng.integers(2, 4))
      return ('pick', [float(w) for w in rng.uniform(0.1, 1.0, ways)], [grow(depth + 1, reads) for _ in range(ways)])
    if kind == 'write':
      return ('write', bool(rng.integers(2)), int(rng.integers(-2, 3)), int(rng.integers(size_a)), grow(depth + 1, reads))
    return ('end',)

  tree = grow(0, 0)

  def body(tape):
    node = tree
    while node[0] != 'end':
      if node[0] == 'read':
        node = node[3][tape.get(node[1], node[2])]
      elif node[0] == 'pick':
        node = node[2][tape.choose(node[1])]
      else:
        tape.set(node[1], node[2], node[3])
        node = node[4]
  return body


@pytest.mark.parametrize('seed', range(32))
def test_random_programs_match_the_oracle(mt, oracle, seed):
  """Random decision trees over the reference's three primitives (tape-get, tape-set!, choose,
  gambit_macros.scm:99-125) - reads after writes, writes to cells never read, several writes to one
  cell, choices before and after reads - through register_program, for windows shorter and longer than
  the cells a program touches: dy/dt against the oracle interpreting the same tree, both modes."""
  from chemical_kinetics_and_program_execution_b200 import programs
  rng = numpy.random.default_rng(5000 + seed)
  size_a = int(rng.integers(2, 4))
  body = random_program(seed, size_a)
  tag = f'fuzz-program-{seed}'
  mt.register_program(tag, size_a, body)
  oracle.register_program(tag, size_a, programs.trace(body, size_a))
  for cl_k in sorted(set(min(int(k), 6 if size_a == 3 else 7) for k in rng.integers(1, 8, size=3))):
    f = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)
    full = configs.markov_table(size_a, cl_k, seed)
    sparse = full * (rng.random(full.shape) > 0.35)
    sparse = sparse / sparse.sum() if sparse.sum() > 0 else full
    for p in (full, sparse):
      got = f(p, 0.0)
      src, dst, w = oracle.terms(tag, cl_k, p, mode=oracle.MERGED)
      gross = numpy.zeros(p.size)
      numpy.add.at(gross, src, abs(w))
      numpy.add.at(gross, dst, abs(w))
      # deep trees put up to 3 * 10^4 terms on one state (seed 7, k = 6): the rounding of a sum in another
      # order grows like the square root of their number (the oracle's own two modes differ by 4.6e-14 of
      # the gross flux there), so the tolerance does beyond 100 terms per state
      scale = max(1.0, (len(w) / p.size / 100.0) ** 0.5)
      assert_rhs_close(got, oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED), gross, tolerance=1e-14 * scale)
      if len(w) <= 3_000_000:  # the literal recursion takes half a minute beyond
        assert_rhs_close(got, oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.LITERAL), gross, tolerance=3e-14 * scale)
      assert abs(got.sum()) <= 1e-14 * scale * gross.sum() + 1e-300
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)


def test_graph_replay_is_bit_identical(mt, device, p0_fixtures):
  """Small problems replay the weight kernels from a CUDA graph captured per input pointer (host
  entry point, stepper, any non-default stream); launching them one by one gives the same bits."""
  import torch
  tag, size_a, cl_k = 'ex4-chemical-turing', 9, 5
  f = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)
  model = device.DeviceModel(tag, cl_k)
  tables = [configs.markov_table(size_a, cl_k, s) for s in (1, 2, 3)]
  p0 = dense(p0_fixtures['ex4_a_idx'], p0_fixtures['ex4_a_val'], 9 ** 5)
  kw = dict(tag=tag, size_a=size_a, cl_k=cl_k, p0=p0, ts=numpy.linspace(0, 30.0, 7), rtol=1e-10, atol=1e-12,
            want_stats=True)
  # more distinct input pointers than graphs are kept, each used twice (capture, then replay)
  bufs = [torch.from_numpy(tables[i % 3]).cuda().clone() for i in range(24)]
  side = torch.cuda.Stream()
  torch.cuda.synchronize()
  results = []
  model.set_option('fused_small', 0)  # the single-launch kernel would bypass the graphs this test is about
  for flag in (1, 0):
    model.set_option('graphs', flag)
    host = [f(p, 0.0) for p in tables for _ in range(2)]
    with torch.cuda.stream(side):
      outs = [model.rhs(b).cpu().numpy() for _ in range(2) for b in bufs]
    results.append((host + outs, mt.ode_integrate_device(**kw)))
  model.set_option('graphs', 1)
  model.set_option('fused_small', 1)
  for x, y in zip(results[0][0], results[1][0]):
    assert numpy.array_equal(x, y)
  assert results[0][1][1] == results[1][1][1] and numpy.array_equal(results[0][1][0], results[1][1][0])
  assert numpy.array_equal(results[0][0][0], results[0][0][6])  # host path and device path agree as well


@pytest.mark.parametrize('tag,size_a,cl_k', [('ex2-ferromagnetic-chain', 2, 3), ('ex2-ferromagnetic-chain', 2, 7),
                                             ('ex1-radioactive-decay', 2, 1), ('ex3-copolymerization', 4, 6),
                                             ('ex5-msrtf-machine', 5, 5), ('ex4-chemical-turing', 9, 5),
                                             ('ex4var2-chemical-turing', 10, 4)])
def test_single_launch_right_hand_side_is_bit_identical(mt, device, oracle, tag, size_a, cl_k):
  """Small problems evaluate the whole right-hand side in one launch (one thread block or one
  cluster of thread blocks running the phases of the separate kernels over virtual blocks,
  engine.cu fused_rhs_kernel).  Every cluster size gives the bits of the multi-launch path - for
  dy/dt, for the node weights, through the host entry point and through the stepper, whose stage
  update rides on the product phase - and dy/dt matches the oracle."""
  import torch
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  model = device.DeviceModel(tag, cl_k)
  assert (model.info['launches_per_rhs'] == 1) == (model.info['n_nodes'] + model.info['nnz'] + size_a ** cl_k <= 8192)
  tables = [configs.markov_table(size_a, cl_k, 7), configs.dirichlet_product_table(size_a, cl_k, 8)]
  f = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)
  p0 = tables[0]
  kw = dict(tag=tag, size_a=size_a, cl_k=cl_k, p0=p0, ts=numpy.linspace(0, 2.0, 5), rtol=1e-9, atol=1e-11, want_stats=True)
  model.set_option('fused_small', 0)
  assert model.info['launches_per_rhs'] > 1
  want = [model.rhs(torch.from_numpy(p).cuda()).cpu().numpy() for p in tables]
  want_w = model.node_weights()
  want_run = mt.ode_integrate_device(**kw)
  for p, w in zip(tables, want):
    assert_rhs_close(w, oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED), gross_flux(oracle, tag, cl_k, p))
  model.set_option('fused_small', 1)
  for cluster in (1, 2, 8, 16):
    model.set_option('fused_cluster', cluster)
    for p, w in zip(tables, want):
      assert numpy.array_equal(model.rhs(torch.from_numpy(p).cuda()).cpu().numpy(), w), cluster
      assert numpy.array_equal(f(p, 0.0), w), cluster
    assert numpy.array_equal(model.node_weights(), want_w), cluster
    run = mt.ode_integrate_device(**kw)
    assert run[1] == want_run[1] and numpy.array_equal(run[0], want_run[0]), cluster
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)


def test_device_rhs_equals_host_rhs(mt, device):
  import torch
  p = configs.markov_table(5, 5, 1)
  host = mt.get_dy_dt(tag='ex5-msrtf-machine', size_a=5, cl_k=5)(p, 0.0)
  model = device.DeviceModel('ex5-msrtf-machine', 5)
  out = model.rhs(torch.from_numpy(p).cuda())
  assert numpy.array_equal(out.cpu().numpy(), host)  # same kernels, deterministic order


def test_flux_row_ranges_and_gather_depths(mt, device):
  """Row blocks that do not start at slice boundaries, every gather depth of the product kernel
  and every load batching of the level kernel give the same bits as one full evaluation."""
  import torch
  p = torch.from_numpy(configs.markov_table(9, 5, 2)).cuda()
  model = device.DeviceModel('ex4-chemical-turing', 5)
  assert model.info['flux_format'] == 1 and model.info['nnz_stored'] < model.info['nnz']
  assert model.info['interleaved_levels'] > 0  # 24 leaf worlds: their chains share the table reads
  want = model.rhs(p).cpu().numpy()
  n = model.n_states
  out = torch.full((n,), float('nan'), dtype=torch.float64, device='cuda')
  model.weights(p)
  cuts = [0, 1, 31, 32, 1000, 1057, n // 2 + 5, n - 1, n]
  for lo, hi in zip(cuts[:-1], cuts[1:]):
    model.flux_rows(out, lo, hi)
  assert numpy.array_equal(out.cpu().numpy(), want)
  guard = torch.full((n,), 7.0, dtype=torch.float64, device='cuda')
  model.flux_rows(guard, 40, 50)  # nothing outside the range is written
  g = guard.cpu().numpy()
  assert (g[:40] == 7.0).all() and (g[50:] == 7.0).all() and numpy.array_equal(g[40:50], want[40:50])
  for key, values in (('flux_unroll', (2, 3, 4, 6, 8)), ('level_unroll', (1, 2, 4, 5, 8)), ('interleave_seeds', (0, 1)), ('ratio_table', (0, 1)),
                      ('plane_kernel', (0, 1)), ('fuse_marginal_ratio', (0, 1))):
    keep = model.info.get(key, None)
    for v in values:
      model.set_option(key, v)
      assert numpy.array_equal(model.rhs(p).cpu().numpy(), want), (key, v)
    if keep:
      model.set_option(key, keep)
  mt.u_lib.tapes_release_model(b'ex4-chemical-turing', 5)  # later tests get a model with default options


@pytest.mark.parametrize('size_a,cl_k,n_rules', [(10, 5, 6), (4, 8, 6), (2, 13, 4), (3, 9, 5), (10, 6, 8)])
def test_plane_kernel_and_left_ratio_table_are_bit_identical(mt, device, oracle, monkeypatch, size_a, cl_k, n_rules):
  """Regular blocks of 256 prefix groups are evaluated by plane_kernel (one 32-byte record per block,
  all loads of a thread independent) and left children with a full window read the left ratio table:
  same operands, same operations, same order of additions as the general level kernel, so node
  weights and dy/dt keep their bits; and they match the oracle."""
  import torch
  rules = configs.random_rule_set(size_a, n_rules, seed=size_a + cl_k)
  tag = f'plane-{size_a}-{cl_k}'
  mt.register_rule_set(tag, size_a, rules)
  oracle.register_rules(tag, size_a, rules)
  p_host = configs.markov_table(size_a, cl_k, 3)
  p = torch.from_numpy(p_host).cuda()
  monkeypatch.setenv('TAPES_RATIO_LEFT', '1')  # the left table is an option (it does not pay at the bench size)
  model = device.DeviceModel(tag, cl_k)
  assert model.info['plane_groups'] > 0 and model.info['ratio_tables'] == 2, model.info
  model.set_option('fused_small', 0)  # the smaller cases here would otherwise take the single-launch kernel
  with_planes, weights = model.rhs(p).cpu().numpy(), model.node_weights()
  assert_rhs_close(with_planes, oracle.compute_dy_dt(tag, cl_k, p_host, mode=oracle.MERGED), gross_flux(oracle, tag, cl_k, p_host))
  launches = model.info['launches_per_rhs']
  model.set_option('plane_kernel', 0)
  assert model.info['launches_per_rhs'] < launches
  assert numpy.array_equal(model.rhs(p).cpu().numpy(), with_planes)
  assert numpy.array_equal(model.node_weights(), weights)
  model.set_option('ratio_table', 0)  # neither table: every node divides
  assert numpy.array_equal(model.rhs(p).cpu().numpy(), with_planes)
  assert numpy.array_equal(model.node_weights(), weights)
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  monkeypatch.delenv('TAPES_RATIO_LEFT')  # the default: right table only
  monkeypatch.setenv('TAPES_MATERIALIZE_RIGHT', '1')  # and the weights of right children written per step
  model = device.DeviceModel(tag, cl_k)
  model.set_option('fused_small', 0)
  assert model.info['ratio_tables'] == 1 and model.info['materialize_right'] == 1
  assert numpy.array_equal(model.rhs(p).cpu().numpy(), with_planes)
  assert numpy.array_equal(model.node_weights(), weights)
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)


@pytest.mark.parametrize('tag,size_a,cl_k', [('ex4-chemical-turing', 9, 5), ('ex3-copolymerization', 4, 7),
                                             ('synthetic', 10, 5), ('synthetic', 2, 12)])
def test_prefix_table_sizing_does_not_change_the_structure(mt, device, monkeypatch, tag, size_a, cl_k):
  """The table that removes duplicate right-chain prefixes during the build is sized on an estimate
  (the nodes that start right chains plus 1.5 times the previous level's prefixes); when the estimate
  is too small the pass starts over with the safe size.  The structure - and so dy/dt and the node weights, bit for bit -
  must not depend on which happened: built with the safe size throughout, with the guess, and
  with a table that is too small at every level (every level retried)."""
  import torch
  if tag == 'synthetic':
    tag = f'hash-{size_a}-{cl_k}'
    mt.register_rule_set(tag, size_a, configs.random_rule_set(size_a, 6, seed=cl_k))
  p = torch.from_numpy(configs.markov_table(size_a, cl_k, 11)).cuda()
  results = {}
  for mode in ('0', '1', '2', 'written-out'):
    # 'written-out': levels of right children only are materialised like the others instead of being
    # read off their groups (Frontier::right_only) - again the same structure
    monkeypatch.setenv('TAPES_HASH_GUESS', '1' if mode == 'written-out' else mode)
    monkeypatch.setenv('TAPES_VIRTUAL_RIGHT_LEVELS', '0' if mode == 'written-out' else '1')
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)
    model = device.DeviceModel(tag, cl_k)
    info = model.info
    results[mode] = (model.rhs(p).cpu().numpy(), model.node_weights(), info['n_nodes'], info['nnz'], info['hash_unique'])
    if mode == '0':
      assert info['hash_retries'] == 0
    if mode == '2' and info['hash_unique'] > 1024 * info['n_levels']:  # some level has more prefixes than 1024 slots
      assert info['hash_retries'] > 0, info
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  for mode in ('1', '2', 'written-out'):
    assert results[mode][2:] == results['0'][2:], mode
    assert numpy.array_equal(results[mode][0], results['0'][0]), mode
    assert numpy.array_equal(results[mode][1], results['0'][1]), mode


@pytest.mark.parametrize('tag,size_a,cl_k', [('ex4-chemical-turing', 9, 5), ('ex5-msrtf-machine', 5, 5),
                                             ('ex3-copolymerization', 4, 7), ('ex2-ferromagnetic-chain', 2, 7),
                                             ('synthetic', 10, 5), ('synthetic', 3, 6)])
def test_fused_right_chain_is_bit_identical(mt, device, monkeypatch, tag, size_a, cl_k):
  """Letting a prefix group evaluate the right children it sums (instead of gathering them after
  the previous level wrote them) changes neither a node weight nor dy/dt by a single bit."""
  import torch
  if tag == 'synthetic':
    tag = f'fuse-test-{size_a}'
    mt.register_rule_set(tag, size_a, configs.random_rule_set(size_a, 9, seed=4))
  p = torch.from_numpy(configs.markov_table(size_a, cl_k, 8)).cuda()
  got = {}
  for fuse in ('0', '1'):
    monkeypatch.setenv('TAPES_LEVEL_FUSE', fuse)
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)
    model = device.DeviceModel(tag, cl_k)
    dy = model.rhs(p).cpu().numpy()
    got[fuse] = (dy, model.node_weights(), dict(model.info))
  monkeypatch.delenv('TAPES_LEVEL_FUSE')
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  assert got['0'][2]['owned_parents'] == 0 and got['0'][2]['deferred_groups'] == 0
  assert got['1'][2]['owned_parents'] == got['1'][2]['deferred_groups'] * size_a
  if cl_k >= 5 and size_a >= 3:
    assert got['1'][2]['owned_parents'] > 0
  assert numpy.array_equal(got['0'][1], got['1'][1])
  assert numpy.array_equal(got['0'][0], got['1'][0])


def test_long_rows(mt, device, oracle):
  """States with more than 128 flux entries (many rules on a small alphabet): runs are only
  tracked for the first 128 entries of a row, the rest goes to columns; term set and dy/dt must
  still match the oracle."""
  import torch
  size_a, cl_k, tag = 2, 7, 'long-rows'
  rules = configs.random_rule_set(size_a, 70, seed=21)
  mt.register_rule_set(tag, size_a, rules)
  oracle.register_rules(tag, size_a, rules)
  p = configs.markov_table(size_a, cl_k, 17)
  model = device.DeviceModel(tag, cl_k)
  row_ptr, entries = model.csr()
  assert numpy.diff(row_ptr).max() > 128
  for lo, hi in zip(row_ptr[:-1], row_ptr[1:]):
    assert (numpy.diff(entries[lo:hi].astype(numpy.int64)) > 0).all()  # canonical order survives the round trip
  got = model.rhs(torch.from_numpy(p).cuda()).cpu().numpy()
  want = oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED)
  assert abs(got - want).max() <= 1e-13 * abs(want).max()
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)


def test_explicit_parent_lists_are_bit_identical(mt, device, monkeypatch):
  """The general form of a level (explicit parent lists) and the usual one (arithmetic
  progressions of parent ids) evaluate the same sums in the same order."""
  import torch
  tag, size_a, cl_k = 'ex4-chemical-turing', 9, 5
  p = torch.from_numpy(configs.markov_table(size_a, cl_k, 11)).cuda()
  got = {}
  for keep in (False, True):
    if keep:
      monkeypatch.setenv('TAPES_KEEP_PARENT_LISTS', '1')
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)
    model = device.DeviceModel(tag, cl_k)
    got[keep] = (model.rhs(p).cpu().numpy(), model.node_weights(), dict(model.info))
  monkeypatch.delenv('TAPES_KEEP_PARENT_LISTS')
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  assert got[False][2]['irregular_levels'] == 0 and got[True][2]['irregular_levels'] > 0
  assert got[True][2]['owned_parents'] == 0
  assert numpy.array_equal(got[False][1], got[True][1])
  assert numpy.array_equal(got[False][0], got[True][0])


@pytest.mark.parametrize('tag,size_a,cl_k,n_parts', [('ex4-chemical-turing', 9, 4, 3),
                                                     ('ex5-msrtf-machine', 5, 5, 8),
                                                     ('ex3-copolymerization', 4, 6, 5),
                                                     ('ex1-radioactive-decay', 2, 4, 2)])
def test_model_parts_sum_to_the_whole(mt, device, oracle, tag, size_a, cl_k, n_parts):
  """The shares of a problem that ranks evaluate together (tapes_model_part; any registered problem,
  flux rules dealt by cost): every flux term lands in exactly one part, with the bits it has in
  the whole problem, and the parts' dy/dt add up to the whole (here all parts on one GPU)."""
  import torch
  p = configs.markov_table(size_a, cl_k, 23)
  d_p = torch.from_numpy(p).cuda()
  whole = device.DeviceModel(tag, cl_k)
  want = whole.rhs(d_p).cpu().numpy()
  torch.cuda.synchronize()
  wsrc, wdst, ww = whole.terms()
  total = numpy.zeros_like(want)
  got_terms = []
  n_rules = 0
  for part in range(n_parts):
    share = device.DeviceModel(tag, cl_k, part=(part, n_parts))
    assert share.handle != whole.handle
    total += share.rhs(d_p).cpu().numpy()
    torch.cuda.synchronize()
    got_terms.append(share.terms())
    n_rules += share.info['n_flux_rules']
  assert n_rules == whole.info['n_flux_rules']
  if tag == 'ex1-radioactive-decay':
    assert got_terms[1][0].size == 0  # more parts than rules: an empty share is a valid model
  src, dst, w = (numpy.concatenate([t[i] for t in got_terms]) for i in range(3))
  key = lambda s, d, x: sorted(zip(s.tolist(), d.tolist(), x.tolist()))
  assert key(src, dst, w) == key(wsrc, wdst, ww)  # same terms, same weights, bit for bit
  assert_rhs_close(total, want, gross_flux(oracle, tag, cl_k, p))
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)


@pytest.mark.parametrize('tag,size_a,cl_k,limit', [('ex4-chemical-turing', 9, 4, 9000),
                                                   ('ex5-msrtf-machine', 5, 5, 40000),
                                                   ('ex3-copolymerization', 4, 6, 1)])
def test_composite_model_matches_single_structure(mt, device, oracle, monkeypatch, tag, size_a, cl_k, limit):
  """A forest too large for the 31-bit node ids of one structure is built as several structures
  over disjoint rule shares and evaluated one after the other (here forced by a small term limit):
  same sizes in total, dy/dt equal to the single structure within rounding, and every entry point
  (device, host buffers, row ranges, fused and unfused stepper) consistent with the others bit for
  bit."""
  import torch
  p = configs.markov_table(size_a, cl_k, 29)
  d_p = torch.from_numpy(p).cuda()
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  single = device.DeviceModel(tag, cl_k)
  assert single.info['structures'] == 1
  want = single.rhs(d_p).cpu().numpy()
  sizes = {key: single.info[key] for key in ('n_nodes', 'nnz', 'n_terms', 'n_flux_rules', 'n_states')}
  kw = dict(tag=tag, size_a=size_a, cl_k=cl_k, p0=p, ts=numpy.linspace(0, 2.0, 5), rtol=1e-9, atol=1e-12,
            want_stats=True)
  want_run = mt.ode_integrate_device(**kw)
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  monkeypatch.setenv('TAPES_MAX_PART_TERMS', str(limit))
  try:
    model = device.DeviceModel(tag, cl_k)
    assert model.info['structures'] > 1
    if limit == 1:
      assert model.info['structures'] == model.info['n_flux_rules']  # never more structures than rules
    assert {key: model.info[key] for key in sizes} == sizes
    got = model.rhs(d_p).cpu().numpy()
    assert_rhs_close(got, want, gross_flux(oracle, tag, cl_k, p))
    assert numpy.array_equal(mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)(p, 0.0), got)
    out = torch.full((model.n_states,), float('nan'), dtype=torch.float64, device='cuda')
    model.weights(d_p)
    cuts = [0, 33, model.n_states // 2 + 1, model.n_states]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
      model.flux_rows(out, lo, hi)
    assert numpy.array_equal(out.cpu().numpy(), got)
    assert model.rhs_profile(d_p, out).shape == (3,) and numpy.array_equal(out.cpu().numpy(), got)
    with pytest.raises(RuntimeError):
      model.csr()
    runs = []
    for flag in ('1', '0'):  # the stage update rides on the last structure's product
      monkeypatch.setenv('TAPES_RK_FUSED', flag)
      runs.append(mt.ode_integrate_device(**kw))
    assert runs[0][1] == runs[1][1] == want_run[1]
    assert numpy.array_equal(runs[0][0], runs[1][0])
    assert abs(runs[0][0] - want_run[0]).max() <= 1e-12 * abs(want_run[0]).max()
  finally:
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)


def test_oversized_structure_is_split_further(mt, device, oracle, monkeypatch):
  """When the node estimate was too optimistic and a structure overflows its node ids, the build
  starts over with more structures (here the node limit is lowered instead of the forest grown)."""
  import torch
  tag, size_a, cl_k = 'ex4-chemical-turing', 9, 4
  p = configs.markov_table(size_a, cl_k, 41)
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  nodes = device.DeviceModel(tag, cl_k).info['n_nodes']
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  monkeypatch.setenv('TAPES_MAX_NODES', str(nodes // 3))
  try:
    model = device.DeviceModel(tag, cl_k)
    assert model.info['structures'] >= 4 and model.info['n_nodes'] == nodes
    got = model.rhs(torch.from_numpy(p).cuda()).cpu().numpy()
    assert_rhs_close(got, oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED), gross_flux(oracle, tag, cl_k, p))
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)
    monkeypatch.setenv('TAPES_MAX_NODES', '10')  # not even one rule fits: a clean error, nothing half built
    with pytest.raises(RuntimeError, match='2\\^31'):
      device.DeviceModel(tag, cl_k)
  finally:
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)


def test_peer_exchange_with_one_rank(mt, device):
  """The fused product + exchange kernels on a world of one: staging slots, owner sums and the
  result vector reproduce the plain product bit for bit (the N > 1 path is checked on the GPU box
  by scripts/check_multi_gpu.py)."""
  import socket
  import torch
  import torch.distributed as dist
  from chemical_kinetics_and_program_execution_b200 import parallel
  with socket.socket() as sock:
    sock.bind(('127.0.0.1', 0))
    port = sock.getsockname()[1]
  dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{port}', rank=0, world_size=1,
                          device_id=torch.device('cuda', torch.cuda.current_device()))
  try:
    model = device.DeviceModel('ex5-msrtf-machine', 5)  # 3125 states: the last ownership block is ragged
    p = torch.from_numpy(configs.markov_table(5, 5, 13)).cuda()
    want = model.rhs(p).cpu().numpy()
    for rounds in (1, 3):
      peer = parallel.PeerExchangeRhs(model, rounds=rounds)
      assert peer.block % (32 * rounds) == 0 and peer.block >= model.n_states
      for _ in range(2):
        got = peer.rhs_full(p)
      peer.check()
      assert numpy.array_equal(got[:model.n_states].cpu().numpy(), want)
      if rounds == 3:  # the stepper on a group of one is the plain stepper (unfused stage updates)
        kw = dict(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=configs.ex5_p0(5), ts=numpy.linspace(0, 3, 4),
                  rtol=1e-10, atol=1e-12)
        assert numpy.array_equal(mt.ode_integrate_device(peer_group=peer, **kw), mt.ode_integrate_device(**kw))
      peer.close()
  finally:
    dist.destroy_process_group()


def _table_check_reference(spd, eps_mpp=1e-100, iterations=0):
  """The same quantities from the dense construction the reference uses
  (framework/markov_tapes.py:81-130, 155-166), for small tables."""
  import itertools
  size_a, k = spd.shape[0], spd.ndim
  clipped = numpy.clip(spd, eps_mpp, 1)
  mpp = clipped / clipped.sum(axis=-1, keepdims=True)
  m = size_a ** (k - 1)
  ctm = numpy.zeros([size_a] * (2 * (k - 1)))
  for idx in itertools.product(range(size_a), repeat=k):
    ctm[idx[1:] + idx[:-1]] += mpp[idx]
  ctm = ctm.reshape(m, m)
  right, left = spd.sum(axis=-1).ravel(), spd.sum(axis=0).ravel()
  out = dict(total=spd.sum(), marginal_distance=numpy.linalg.norm(right - left),
             stationarity_residual=numpy.linalg.norm(ctm @ left - left))
  if iterations:
    v = numpy.full(m, 1.0 / m)
    for _ in range(iterations):
      v = ctm @ v
    out['power_distance'] = numpy.linalg.norm(v - left)
  return out


def test_check_table_matches_dense_reference(mt, p0_fixtures):
  """Sparse validation of a table (csrc/validate.cu) against the reference's dense transfer
  matrix: shipped starting tables are stationary, a perturbed one is not, and the power iteration
  finds the context marginal of an ergodic table."""
  import torch
  ex5 = configs.ex5_p0(5).reshape([5] * 5)
  markov = configs.markov_table(3, 4, 7).reshape([3] * 4)
  bent = markov.copy()
  bent[0, 1, 2, 0] += 0.01
  bent[2, 2, 1, 1] -= 0.004
  for name, spd, iterations in (('ex5', ex5, 0), ('markov', markov, 64), ('bent', bent, 16)):
    want = _table_check_reference(spd, iterations=iterations)
    got = mt.check_table(spd, max_iterations=iterations, tolerance=0.0)
    for key, value in want.items():
      assert abs(got[key] - value) <= 1e-12 * max(1.0, abs(value)), (name, key, got[key], value)
    on_device = mt.check_table(torch.from_numpy(spd.copy()).cuda().reshape(-1), size_a=spd.shape[0], cl_k=spd.ndim,
                               max_iterations=iterations, tolerance=0.0)
    assert on_device == got
  assert mt.check_table(ex5)['marginal_distance'] <= 1e-15 and mt.check_table(ex5)['stationarity_residual'] <= 1e-15
  ergodic = mt.check_table(markov, max_iterations=2000, tolerance=1e-15)
  assert ergodic['power_distance'] <= 1e-12 and ergodic['iterations'] < 2000
  assert mt.check_table(bent)['marginal_distance'] > 1e-3
  # the reference's own verdicts on the same tables (dense eig) agree
  delta, eigenspace = mt.get_ctm_eigenvalue1_eigenspace(markov)
  assert eigenspace is not None and eigenspace.shape[1] == 1 and delta < 1e-12
  delta, eigenspace = mt.get_ctm_eigenvalue1_eigenspace(bent)
  assert eigenspace is None and abs(delta - mt.check_table(bent)['marginal_distance']) < 1e-12


def test_csr_export_round_trip(mt, device, monkeypatch):
  """The sliced form expands back to the same canonical CSR the plain build keeps, for both
  encoders (full runs only; masked runs from a merge of the 32 rows)."""
  import torch
  tag, cl_k = 'ex3-copolymerization', 6
  results = {}
  for name, env in (('csr', dict(TAPES_FLUX_FORMAT='csr')), ('full', dict(TAPES_RUN_MIN_LANES='32')),
                    ('masked', dict(TAPES_RUN_MIN_LANES='3'))):
    for k_, v_ in env.items():
      monkeypatch.setenv(k_, v_)
    mt.u_lib.tapes_release_model(tag.encode(), cl_k)
    model = device.DeviceModel(tag, cl_k)
    p = torch.from_numpy(configs.markov_table(4, cl_k, 5)).cuda()
    results[name] = (model.csr(), model.rhs(p).cpu().numpy(), dict(model.info))
    for k_ in env:
      monkeypatch.delenv(k_)
  mt.u_lib.tapes_release_model(tag.encode(), cl_k)
  (rp, en), dy, info = results['csr']
  assert info['flux_format'] == 0
  for name in ('full', 'masked'):
    (rp2, en2), dy2, info2 = results[name]
    assert info2['flux_format'] == 1
    assert info2['run_entries'] + info2['column_entries'] == info2['nnz_stored'] < info2['nnz']
    assert numpy.array_equal(rp, rp2) and numpy.array_equal(en, en2)
    scale = abs(dy).max()
    assert abs(dy2 - dy).max() <= 1e-14 * scale
  assert results['masked'][2]['run_entries'] >= results['full'][2]['run_entries']


def test_errors(mt):
  with pytest.raises(ValueError):
    mt.get_dy_dt(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=3)(numpy.ones(5), 0.0)
  with pytest.raises(RuntimeError):
    mt.get_dy_dt(tag='no-such-problem', size_a=2, cl_k=3)(numpy.ones(8) / 8, 0.0)
  with pytest.raises(ValueError):
    mt.ode_integrate(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=3, p0=numpy.ones(8), ts=[0, 1])


def test_ex4_end_points_match_reference(mt, known_answers, p0_fixtures):
  """examples/ex4_chemical_turing.py:98-116 and 150-170 through the drop-in API."""
  for name in ('a', 'b'):
    p0 = dense(p0_fixtures[f'ex4_{name}_idx'], p0_fixtures[f'ex4_{name}_val'], 9 ** 5)
    ys = mt.ode_integrate_ivp(tag='ex4-chemical-turing', size_a=9, cl_k=5, p0=p0,
                              ts=numpy.linspace(0, 2000.0, 2001),
                              ivp_kwargs=dict(rtol=1e-13, atol=1e-13, method='DOP853'))
    assert ys.shape == (2001, 9 ** 5)
    spd = ys[-1].reshape([9] * 5)
    got = [mt.seq_prob(spd, s)[0] for s in known_answers['ex4_observables']]
    for g, w in zip(got, known_answers[f'ex4_p0_{name}_t2000']):
      assert abs(g - w) <= 1e-12 * abs(w), (name, g, w)


def test_ex2_trajectory_matches_oracle(mt, trajectories, p0_fixtures):
  """examples/ex2_ferromagnet_tape.py:74-84 for k = 3..7.

  Through DOP853 on a fixed step sequence (see tests/golden/make_golden.py EX2_FIXED_STEP) the GPU
  and CPU trajectories agree to 1e-12.  Through the shipped odeint/LSODA call (rtol=atol=1e-9)
  they agree to 1e-10: adaptive step control and LSODA's numerically differentiated Jacobian
  amplify last-bit differences of dy/dt - far below the solver's own 1e-9 tolerance, but above
  1e-12 (a 1-ulp change of p0 moves the CPU result by 8e-13 on its own).
  """
  for k in range(3, 8):
    p0 = dense(p0_fixtures[f'ex2_k{k}_idx'], p0_fixtures[f'ex2_k{k}_val'], 2 ** k)
    ys = mt.ode_integrate_ivp(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=k, p0=p0, ts=[0.0, 30.0, 60.0],
                              ivp_kwargs=dict(method='DOP853', rtol=1e-3, atol=1e-6, max_step=0.05,
                                              first_step=0.05))
    want = trajectories[f'ex2_k{k}_dop853_end']
    assert abs(ys[-1] - want).max() <= 1e-12 * abs(want).max()
    ys = mt.ode_integrate(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=k, p0=p0,
                          ts=numpy.linspace(0, 60, 1001), odeint_kwargs=dict(rtol=1e-9, atol=1e-9))
    want = trajectories[f'ex2_k{k}_end']
    assert ys.shape == (1001, 2 ** k)
    assert abs(ys[-1] - want).max() <= 1e-10 * abs(want).max()


def test_ex2_island_probabilities_follow_the_analytic_approximation(mt, p0_fixtures):
  """BASELINE config 1's own check (examples/ex2_ferromagnet_tape.py:112-135, plots only in the
  reference): island probabilities p(0 1^L 0), L = 1..5, of the cl_k = 7 run against the reference's
  analytic approximation (examples/ex2_ferromagnet_analytic.py:39-61, golden produced by importing
  that module: tests/golden/make_golden_round2.py).  The approximation neglects island merging and
  correlations beyond nearest neighbours: the two agree to a few per cent (L <= 4: 5 %; L = 5, whose
  probabilities are ~1e-5 and which the cl_k = 7 closure resolves least: 15 %), at t = 15, 30, 60."""
  import os
  from conftest import GOLDEN
  gold = numpy.load(os.path.join(GOLDEN, 'ex2_analytic.npz'))
  k = 7
  p0 = dense(p0_fixtures[f'ex2_k{k}_idx'], p0_fixtures[f'ex2_k{k}_val'], 2 ** k)
  ts = numpy.linspace(0, 60, 1001)
  assert numpy.array_equal(ts, gold['ts'])
  ys = mt.ode_integrate(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=k, p0=p0, ts=ts,
                        odeint_kwargs=dict(rtol=1e-9, atol=1e-9))
  table = ys.reshape((len(ys),) + (2,) * k)
  for length in (1, 2, 3, 4, 5):
    probs = mt.seq_prob(table, (0, *((1,) * length), 0), num_prefix_indices=1)[0]
    for i in (250, 500, 1000):
      want = gold['islands'][i, length - 1]
      tol = 0.05 if length <= 4 else 0.15
      assert abs(probs[i] - want) <= tol * want, (length, ts[i], probs[i], want)


def test_ex5_trajectory_matches_oracle(mt, trajectories, p0_fixtures):
  p0 = dense(p0_fixtures['ex5_idx'], p0_fixtures['ex5_val'], 5 ** 5)
  ys = mt.ode_integrate_ivp(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=p0,
                            ts=numpy.linspace(0, 500.0, 4001),
                            ivp_kwargs=dict(rtol=1e-13, atol=1e-13, method='DOP853'))
  for key, row in (('ex5_t50', 400), ('ex5_end', 4000)):
    want = trajectories[key]
    assert abs(ys[row] - want).max() <= 1e-12 * abs(want).max()


def test_device_dop853_matches_scipy_dop853(mt, p0_fixtures):
  """The HBM-resident stepper against scipy.integrate.solve_ivp(DOP853) driving the same GPU
  right-hand side: same controller, same step sequence, results within 1e-12."""
  p0 = dense(p0_fixtures['ex5_idx'], p0_fixtures['ex5_val'], 5 ** 5)
  ts = numpy.linspace(0, 40.0, 81)
  kw = dict(rtol=1e-10, atol=1e-12)
  want = mt.ode_integrate_ivp(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=p0, ts=ts,
                              ivp_kwargs=dict(method='DOP853', **kw))
  f = mt.get_dy_dt(tag='ex5-msrtf-machine', size_a=5, cl_k=5)
  sol = scipy.integrate.solve_ivp(lambda t, y: f(y, t), (ts[0], ts[-1]), p0, t_eval=ts, method='DOP853', **kw)
  seqs = [[0], [1, 2], [2, 2, 0, 1, 1]]
  (got, series), stats = mt.ode_integrate_device(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=p0, ts=ts,
                                                 observables=seqs, want_stats=True, **kw)
  assert got.shape == want.shape
  assert stats['nfev'] == sol.nfev
  assert abs(got - want).max() <= 1e-12 * abs(want).max()
  for j, seq in enumerate(seqs):
    ref = mt.seq_prob(got.reshape((len(ts),) + (5,) * 5), seq, num_prefix_indices=1)[0]
    assert abs(series[:, j] - ref).max() <= 1e-14 * abs(ref).max()


def test_device_dop853_reproduces_ex4_reference_end_points(mt, known_answers, p0_fixtures):
  """examples/ex4_chemical_turing.py:150-170 with the table kept in HBM and only the eight
  observables read back."""
  for name in ('a', 'b'):
    p0 = dense(p0_fixtures[f'ex4_{name}_idx'], p0_fixtures[f'ex4_{name}_val'], 9 ** 5)
    series = mt.ode_integrate_device(tag='ex4-chemical-turing', size_a=9, cl_k=5, p0=p0,
                                     ts=numpy.linspace(0, 2000.0, 2001), rtol=1e-13, atol=1e-13,
                                     observables=known_answers['ex4_observables'], return_states=False)
    assert series.shape == (2001, 8)
    for g, w in zip(series[-1], known_answers[f'ex4_p0_{name}_t2000']):
      assert abs(g - w) <= 1e-12 * abs(w), (name, g, w)


def test_device_observables(mt, device):
  import torch
  p = configs.markov_table(4, 6, 5)
  model = device.DeviceModel('ex3-copolymerization', 6)
  seqs = [[1], [0, 0], [3, 1, 0], [0, 1, 2, 3, 0, 1]]
  got = model.observe(torch.from_numpy(p).cuda(), seqs)
  spd = p.reshape([4] * 6)
  for g, seq in zip(got, seqs):
    want = mt.seq_prob(spd, seq)[0]
    assert abs(g - want) <= 1e-14 * abs(want)


def test_device_observables_beyond_the_window_and_entropy(mt, device, p0_fixtures):
  """SURVEY.md section 8(f) rank 1, the rest of it: `seq_prob` for sequences longer than cl_k
  (framework/markov_tapes.py:223-233, Markov extension with the clipped process parameters) and
  `markov_entropy` (markov_tapes.py:178-187) evaluated on the device, against the host functions."""
  import torch
  model = device.DeviceModel('ex3-copolymerization', 6)
  tables = [configs.markov_table(4, 6, 5), configs.dirichlet_product_table(4, 6, 9),
            dense(p0_fixtures['ex3_k6_idx'], p0_fixtures['ex3_k6_val'], 4 ** 6)]  # the last has zeros: clips matter
  seqs = [[0, 1, 2, 3, 0, 1, 2], [1, 3, 1, 2, 0, 0, 1, 2, 3, 3, 0], [2], [0, 0, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0]]
  for p in tables:
    d_p = torch.from_numpy(p).cuda()
    got = model.observe(d_p, seqs)
    spd = p.reshape([4] * 6)
    for g, seq in zip(got, seqs):
      want = float(mt.seq_prob(spd, seq)[0])
      assert abs(g - want) <= 1e-13 * abs(want) + 1e-300, (seq, g, want)
    want_h = float(mt.markov_entropy(spd))
    assert abs(model.entropy(d_p) - want_h) <= 1e-13 * abs(want_h), (model.entropy(d_p), want_h)
  # eps is honoured (framework/markov_tapes.py:101)
  p = tables[2]
  got = model.observe(torch.from_numpy(p).cuda(), [[0, 0, 0, 0, 0, 0, 1, 3]], eps=1e-30)[0]
  want = float(mt.seq_prob(p.reshape([4] * 6), [0, 0, 0, 0, 0, 0, 1, 3], eps=1e-30)[0])
  assert abs(got - want) <= 1e-13 * abs(want)
  with pytest.raises(RuntimeError, match='alphabet'):
    model.observe(torch.from_numpy(p).cuda(), [[0, 7]])


def test_observable_sums_use_the_whole_grid(mt, device):
  """A short sequence at a large table is a strided sum over every sector of the table; it is now
  evaluated by a grid of blocks with a fixed order of additions (was one block per observable)."""
  import time
  import torch
  rules = configs.random_rule_set(10, 2, seed=3)
  mt.register_rule_set('observe-at-scale', 10, rules)
  model = device.DeviceModel('observe-at-scale', 7)  # 10^7 states
  p = configs.dirichlet_product_table(10, 7, 4)
  d_p = torch.from_numpy(p).cuda()
  seqs = [[3], [9, 0], [1, 2, 3, 4, 5, 6, 7], [4, 4, 4]]
  got = model.observe(d_p, seqs)
  spd = p.reshape([10] * 7)
  for g, seq in zip(got, seqs):
    want = float(mt.seq_prob(spd, seq)[0])
    assert abs(g - want) <= 1e-13 * abs(want)
  assert numpy.array_equal(got, model.observe(d_p, seqs))  # reproducible
  t0 = time.perf_counter()
  for _ in range(20):
    model.observe(d_p, seqs)
  print(f'4 observables at 10^7 states: {(time.perf_counter() - t0) / 20 * 1e6:.0f} us per call')
  mt.u_lib.tapes_release_model(b'observe-at-scale', 7)


def test_stepper_returns_long_sequences_and_entropy(mt, p0_fixtures):
  p0 = dense(p0_fixtures['ex5_idx'], p0_fixtures['ex5_val'], 5 ** 5)
  ts = numpy.linspace(0, 5.0, 6)
  seqs = [[0, 1], [0, 1, 2, 0, 1, 2, 0], [4, 4, 4, 4, 4, 4]]
  states, series, entropy = mt.ode_integrate_device(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=p0, ts=ts, rtol=1e-10,
                                                    atol=1e-12, observables=seqs, entropy=True)
  for i in range(ts.size):
    spd = states[i].reshape([5] * 5)
    for j, seq in enumerate(seqs):
      want = float(mt.seq_prob(spd, seq)[0])
      assert abs(series[i, j] - want) <= 1e-13 * abs(want) + 1e-300
    want_h = float(mt.markov_entropy(spd))
    assert abs(entropy[i] - want_h) <= 1e-13 * abs(want_h)


def test_fused_stage_update_is_bit_identical(mt, p0_fixtures, monkeypatch):
  """The Runge-Kutta stage update fused into the product kernel keeps the term order of the
  separate kernel, so both integrators produce the same bits."""
  p0 = dense(p0_fixtures['ex5_idx'], p0_fixtures['ex5_val'], 5 ** 5)
  ts = numpy.linspace(0, 20.0, 11)
  runs = []
  for flag in ('1', '0'):
    monkeypatch.setenv('TAPES_RK_FUSED', flag)
    runs.append(mt.ode_integrate_device(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=p0, ts=ts,
                                        rtol=1e-10, atol=1e-12, want_stats=True))
  assert runs[0][1] == runs[1][1]
  assert numpy.array_equal(runs[0][0], runs[1][0])


def test_reregistering_a_tag_invalidates_handles_instead_of_freeing_under_them(mt, device):
  """registering a tag again (or releasing the model) used to free the structure under
  live DeviceModel handles and steppers.  Now the old handle is refused with a message, a stepper
  created before keeps working on the structure it was created with, and a fresh handle works."""
  import torch
  rules = configs.random_rule_set(3, 4, seed=11)
  tag = 'reregistered-rule-set'
  mt.register_rule_set(tag, 3, rules)
  old = device.DeviceModel(tag, 4)
  p = torch.from_numpy(configs.dirichlet_product_table(3, 4, 5)).cuda()
  want = old.rhs(p).cpu().numpy()
  solver_states = mt.ode_integrate_device(tag=tag, size_a=3, cl_k=4, p0=p.cpu().numpy(), ts=[0.0, 0.5],
                                          rtol=1e-10, atol=1e-12)
  mt.register_rule_set(tag, 3, rules)  # same rules: same answers from a new structure
  with pytest.raises(RuntimeError, match='invalidated'):
    old.rhs(p)
  with pytest.raises(RuntimeError, match='invalidated'):
    old.node_weights()
  fresh = device.DeviceModel(tag, 4)
  assert numpy.array_equal(fresh.rhs(p).cpu().numpy(), want)
  again = mt.ode_integrate_device(tag=tag, size_a=3, cl_k=4, p0=p.cpu().numpy(), ts=[0.0, 0.5], rtol=1e-10, atol=1e-12)
  assert numpy.array_equal(again, solver_states)
  assert mt.u_lib.tapes_release_model(tag.encode(), 4) == 0
  with pytest.raises(RuntimeError, match='invalidated'):
    fresh.rhs(p)


def test_two_streams_on_one_model_do_not_overlap_on_its_scratch(mt, device):
  """the weights of a right-hand side live in per-model scratch and the caller picks the
  stream.  Calls on different streams are ordered by the library (engine.h Model::busy): results of
  interleaved asynchronous calls equal those of synchronised ones."""
  import torch
  tag, size_a, cl_k = 'ex4-chemical-turing', 9, 5
  model = device.DeviceModel(tag, cl_k)
  tables = [torch.from_numpy(configs.markov_table(size_a, cl_k, s)).cuda() for s in range(4)]
  want = []
  for t in tables:
    want.append(model.rhs(t).clone())
    torch.cuda.synchronize()
  f = mt.get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k)
  host = [t.cpu().numpy() for t in tables]
  streams = [torch.cuda.Stream() for _ in range(3)]
  torch.cuda.synchronize()
  for rep in range(20):
    outs = []
    for i, t in enumerate(tables):
      with torch.cuda.stream(streams[(rep + i) % 3]):
        outs.append(model.rhs(t))        # asynchronous, caller's stream
      if i == 1:
        got_host = f(host[2], 0.0)        # the model's own stream, right behind an asynchronous call
        assert numpy.array_equal(got_host, want[2].cpu().numpy())
    torch.cuda.synchronize()
    for o, w in zip(outs, want):
      assert torch.equal(o, w)


def test_failure_through_the_reference_binding_is_loud(mt):
  """A failing c_compute_dy_dt on a box WITH a device: the result buffer of the reference's binding
  (zero-filled, framework/markov_tapes.py:279) comes back NaN, not zero."""
  from test_abi import reference_binding, reference_dy_dt
  u_lib = reference_binding()
  out = reference_dy_dt(u_lib, 'no-such-problem', 3, numpy.full(8, 0.125))
  assert numpy.isnan(out[0])
  # cl_k = 0 is refused by the engine after the tag was found: the table size (A^0 = 1) is known
  out = reference_dy_dt(u_lib, 'ex1-radioactive-decay', 0, numpy.full(1, 1.0))
  assert numpy.isnan(out).all()
  mt.u_lib.tapes_clear_error()
  # and the next good call is not disturbed by the message the failures left behind
  good = mt.get_dy_dt(tag='__canary_problem_radioactive_decay', size_a=2, cl_k=3)(numpy.full(8, 0.125), 0.0)
  assert good.tolist() == [0.375, 0.125, 0.125, -0.125, 0.125, -0.125, -0.125, -0.375]
