"""CPU tests (gloo, world_size 2) of the multi-GPU plumbing: rule dealing, flux reduce-scatter
and table all-gather reproduce the unsplit right-hand side.  The local right-hand side is played
by the CPU oracle here; on the GPU box bench.py plugs in the CUDA path."""

import os
import socket

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chemical_kinetics_and_program_execution_b200 import configs, parallel

SIZE_A, CL_K, N_RULES = 5, 4, 7


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, result_path):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from oracle import oracle
  rules = configs.random_rule_set(SIZE_A, N_RULES, seed=6)
  local = parallel.split_rule_set(rules, world, rank)
  tag = f'par-test-{rank}'
  oracle.register_rules(tag, SIZE_A, local)
  n = SIZE_A ** CL_K

  def local_rhs(p_full, out_full):
    out_full.copy_(torch.from_numpy(oracle.compute_dy_dt(tag, CL_K, p_full.numpy(), mode=oracle.MERGED)))

  sharded = parallel.ShardedRhs(local_rhs, n, device='cpu')
  p = torch.zeros(sharded.padded, dtype=torch.float64)
  p[:n] = torch.from_numpy(configs.markov_table(SIZE_A, CL_K, 3))
  out = torch.zeros_like(p)
  sharded.rhs_full(p, out)

  # the chunked, overlapped exchange must give the same vector
  cache = {}

  def local_weights(p_full):
    cache['dy'] = torch.from_numpy(oracle.compute_dy_dt(tag, CL_K, p_full.numpy(), mode=oracle.MERGED))

  def local_flux_rows(out_partial, lo, hi):
    out_partial[lo:hi] = cache['dy'][lo:hi]

  over = parallel.OverlappedRhs(local_weights, local_flux_rows, n, chunks=3, device='cpu')
  p2 = torch.zeros(over.padded, dtype=torch.float64)
  p2[:n] = p[:n]
  out2 = torch.zeros_like(p2)
  over.rhs_full(p2, out2)
  assert torch.equal(out2[:n], out[:n])
  assert float(out2[n:].abs().sum()) == 0.0

  # the default exchange of bench.py: all-reduce in row blocks that start at multiples of 32
  summed = parallel.OverlappedAllReduceRhs(local_weights, local_flux_rows, n, chunks=3)
  assert [lo % 32 for lo, _ in summed.bounds] == [0] * len(summed.bounds)
  assert summed.bounds[0][0] == 0 and summed.bounds[-1][1] == n
  out3 = torch.zeros(n, dtype=torch.float64)
  summed.rhs_full(p[:n].clone(), out3)
  assert torch.equal(out3, out[:n])
  if rank == 0:
    numpy.save(result_path, out[:n].numpy())
  dist.destroy_process_group()


def test_split_rule_set_keeps_every_rule_once():
  rules = configs.random_rule_set(SIZE_A, N_RULES, seed=6)
  seen = []
  for r in range(3):
    part = parallel.split_rule_set(rules, 3, r)
    assert abs(part['select_weight'].sum() - rules['select_weight'].sum()) < 1e-12
    seen += part['rate'][:-1].tolist()  # last one is the inert rule
    assert (part['pattern'][-1] == part['repl'][-1]).all()
  assert sorted(seen) == sorted(rules['rate'].tolist())


def test_balanced_dealing_matches_term_counts(oracle):
  rules = configs.random_rule_set(SIZE_A, N_RULES, seed=6)
  costs = parallel.rule_costs(rules, SIZE_A, CL_K)
  p = configs.dirichlet_product_table(SIZE_A, CL_K, 1)
  for r in range(N_RULES):  # the cost model is exact: it counts the oracle's flux terms
    one = {key: numpy.asarray(val)[r:r + 1] for key, val in rules.items()}
    oracle.register_rules('cost-probe', SIZE_A, one)
    src, _, _ = oracle.terms('cost-probe', CL_K, p, mode=oracle.MERGED)
    assert len(src) == costs[r]
  owner = parallel.deal_rules(costs, 3)
  loads = [costs[owner == g].sum() for g in range(3)]
  assert max(loads) - min(loads) <= costs.max()
  seen = []
  for g in range(3):
    part = parallel.split_rule_set(rules, 3, g, SIZE_A, CL_K)
    seen += part['rate'][:-1].tolist()
  assert sorted(seen) == sorted(rules['rate'].tolist())


def test_rotated_rule_sets_have_equal_cost():
  base = configs.random_rule_set(SIZE_A, N_RULES, seed=6)
  want = parallel.rule_costs(base, SIZE_A, CL_K)
  for shift in (1, 3):
    rot = configs.rotated_rule_set(base, shift, SIZE_A)
    assert (parallel.rule_costs(rot, SIZE_A, CL_K) == want).all()
    assert (rot['pattern'] != base['pattern']).any()
  both = configs.concat_rule_sets([base, configs.rotated_rule_set(base, 1, SIZE_A)])
  part = parallel.take_rules(both, numpy.arange(N_RULES, 2 * N_RULES))
  assert part['select_weight'][-1] == N_RULES and len(part['rate']) == N_RULES + 1


def test_block_bounds_cover_all_states():
  for n, w in ((625, 2), (1000, 8), (7, 4)):
    covered = []
    for r in range(w):
      lo, hi, block = parallel.block_bounds(n, w, r)
      covered += list(range(lo, hi))
      assert hi - lo <= block
    assert covered == list(range(n))


def test_two_rank_rhs_equals_unsplit(tmp_path, oracle):
  result = str(tmp_path / 'dy.npy')
  port = _free_port()
  mp.spawn(_worker, args=(2, port, result), nprocs=2, join=True)
  got = numpy.load(result)
  rules = configs.random_rule_set(SIZE_A, N_RULES, seed=6)
  oracle.register_rules('par-test-full', SIZE_A, rules)
  p = configs.markov_table(SIZE_A, CL_K, 3)
  want = oracle.compute_dy_dt('par-test-full', CL_K, p, mode=oracle.MERGED)
  assert abs(got - want).max() <= 1e-14 * abs(want).max()


def _parts_worker(rank, world, port, result_path):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from chemical_kinetics_and_program_execution_b200 import _lib
  report = {}
  for tag, cl_k in (('ex4-chemical-turing', 5), ('ex5-msrtf-machine', 5), ('ex3-copolymerization', 12)):
    owner, cost = _lib.rule_parts(tag, cl_k, world)
    everyone = [None] * world
    dist.all_gather_object(everyone, (owner.tolist(), cost.tolist()))
    assert all(e == everyone[0] for e in everyone)  # every rank deals the same way without talking
    mine = numpy.nonzero(owner == rank)[0]
    counts = [None] * world
    dist.all_gather_object(counts, (len(mine), float(cost[mine].sum())))
    assert sum(c[0] for c in counts) == len(owner)
    loads = [c[1] for c in counts]
    assert max(loads) <= sum(loads) / world + cost.max()
    report[tag] = loads
  if rank == 0:
    numpy.save(result_path, numpy.array([report[t] for t in sorted(report)]))
  dist.destroy_process_group()


def test_library_deals_registered_problems_consistently_on_two_ranks(tmp_path):
  """tapes_model_part's dealing (host-only part, no GPU): two processes compute it independently
  and must agree on every flux rule's owner, with balanced loads."""
  path = str(tmp_path / 'loads.npy')
  mp.spawn(_parts_worker, args=(2, _free_port(), path), nprocs=2, join=True)
  loads = numpy.load(path)
  assert loads.shape == (3, 2) and (loads > 0).all()
  assert (abs(loads[:, 0] - loads[:, 1]) <= 0.1 * loads.sum(axis=1)).all()
