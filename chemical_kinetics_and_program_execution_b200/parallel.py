"""Multi-GPU evaluation of the master-equation step (one process per GPU, torch.distributed for
set-up and barriers).

The window-extension forests of different flux rules are independent, so the rules are dealt to
the ranks: rank g builds and evaluates only its share over the full state space and produces a
partial dy/dt; the partial results are summed by one of the exchanges below.  The library deals any
registered problem itself (device.DeviceModel(tag, k, part=(rank, world)), tapes_model_part); the
helpers here do the same for rewrite-rule dicts, which the weak-scaling benchmark uses to give every
rank a rotated copy of the same rules.

  PeerExchangeRhs          the default: the exchange runs inside the product kernel over NVLink peer
                           memory (tapes_peer_rhs), no NCCL kernel on the data path
  OverlappedAllReduceRhs   NCCL all-reduce of dy/dt in row blocks, overlapped with the product
  OverlappedRhs, ShardedRhs   NCCL reduce-scatter + all-gather, with and without overlap

The NCCL variants touch the GPU only through torch, so their plumbing is testable with the gloo
backend on CPU (tests/test_parallel.py).
"""

import numpy
import torch
import torch.distributed as dist


def rule_costs(rules, size_a, cl_k):
  """Flux terms each rule generates at (size_a, cl_k): one per length-k window that overlaps a
  changed cell and per assignment of the window cells outside the rule's span."""
  pattern = numpy.asarray(rules['pattern']).reshape(-1, 4)
  repl = numpy.asarray(rules['repl']).reshape(-1, 4)
  costs = []
  for r, span in enumerate(numpy.asarray(rules['span'])):
    changed = [c for c in range(int(span)) if pattern[r, c] != repl[r, c]]
    if not changed:
      costs.append(0.0)
      continue
    total = 0.0
    for start in range(changed[0] - cl_k + 1, changed[-1] + 1):
      inside = max(0, min(start + cl_k, int(span)) - max(start, 0))
      total += float(size_a) ** (cl_k - inside)
    costs.append(total)
  return numpy.array(costs)


def deal_rules(costs, world_size):
  """Longest-processing-time assignment: rules in descending cost to the least loaded rank
  (ties to the lower rank), so every rank computes the same dealing."""
  order = sorted(range(len(costs)), key=lambda r: (-costs[r], r))
  load = [0.0] * world_size
  owner = [0] * len(costs)
  for r in order:
    g = min(range(world_size), key=lambda j: (load[j], j))
    owner[r] = g
    load[g] += costs[r]
  return numpy.array(owner)


def take_rules(rules, mine):
  """The rules with indices `mine` plus one inert rule that carries the selection weight of all
  the others, so that the probability of picking a kept rule is the same as in the full problem
  (gambit_macros.scm:75-86 normalises by the sum of the weights)."""
  n = len(rules['rate'])
  mine = numpy.asarray(mine, dtype=numpy.int64)
  others = numpy.setdiff1d(numpy.arange(n), mine)
  out = {key: numpy.asarray(val)[mine] for key, val in rules.items()}
  if len(others):
    rest = float(numpy.asarray(rules['select_weight'])[others].sum())
    zeros4 = numpy.zeros((1, 4), dtype=numpy.int32)
    inert = dict(tape=[0], span=[1], catalyst=[-1], pattern=zeros4, repl=zeros4, rate=[1.0],
                 select_weight=[rest])
    out = {key: numpy.concatenate([numpy.asarray(out[key]), numpy.asarray(inert[key], dtype=numpy.asarray(out[key]).dtype)])
           for key in out}
  return out


def split_rule_set(rules, world_size, rank, size_a=None, cl_k=None):
  """This rank's share of a rule-set dict (see take_rules).  With size_a and cl_k the rules are
  dealt by estimated cost (balanced), else round-robin."""
  n = len(rules['rate'])
  if size_a is not None and cl_k is not None:
    mine = numpy.nonzero(deal_rules(rule_costs(rules, size_a, cl_k), world_size) == rank)[0]
  else:
    mine = numpy.arange(rank, n, world_size)
  return take_rules(rules, mine)


def block_bounds(n_states, world_size, rank):
  """Contiguous ownership blocks of equal padded size."""
  block = -(-n_states // world_size)
  lo = min(rank * block, n_states)
  return lo, min(lo + block, n_states), block


class ShardedRhs:
  """dy/dt of the full problem from per-rank partial right-hand sides.

  local_rhs(p_full, out_full) must write this rank's partial dy/dt (all states) into out_full.
  """

  def __init__(self, local_rhs, n_states, group=None, device='cuda'):
    self.local_rhs = local_rhs
    self.group = group
    self.world = dist.get_world_size(group)
    self.rank = dist.get_rank(group)
    self.n = n_states
    _, _, self.block = block_bounds(n_states, self.world, self.rank)
    self.padded = self.block * self.world
    self.partial = torch.zeros(self.padded, dtype=torch.float64, device=device)
    self.mine = torch.zeros(self.block, dtype=torch.float64, device=device)

  def owned_flux(self, p_full):
    """This rank's block of the summed dy/dt (length `block`, zero-padded at the end)."""
    self.local_rhs(p_full[:self.n], self.partial[:self.n])
    dist.reduce_scatter_tensor(self.mine, self.partial, op=dist.ReduceOp.SUM, group=self.group)
    return self.mine

  def gather(self, block_values, out_full):
    """All-gathers per-rank blocks into the padded full vector."""
    dist.all_gather_into_tensor(out_full, block_values, group=self.group)
    return out_full

  def rhs_full(self, p_full, out_full):
    return self.gather(self.owned_flux(p_full), out_full)


class OverlappedAllReduceRhs:
  """dy/dt of the full problem on every rank: partial flux summed by all-reduce, row block by row
  block, so that the exchange of one block runs while the product of the next is computed.

  local_weights(p) evaluates everything that depends on p; local_flux_rows(out, lo, hi) writes this
  rank's partial dy/dt for the states lo <= i < hi.  Blocks start at multiples of 32 states (the
  slices of the product kernel).
  """

  def __init__(self, local_weights, local_flux_rows, n_states, chunks=8, group=None):
    self.local_weights = local_weights
    self.local_flux_rows = local_flux_rows
    self.group = group
    self.n = n_states
    chunks = max(1, int(chunks))
    per = -(-n_states // chunks)
    per = -(-per // 32) * 32
    self.bounds = [(lo, min(lo + per, n_states)) for lo in range(0, n_states, per)]

  def rhs_full(self, p_full, out_full):
    """p_full, out_full: vectors of at least n_states doubles; out_full[:n] receives the sum."""
    self.local_weights(p_full[:self.n])
    pending = []
    for lo, hi in self.bounds:
      self.local_flux_rows(out_full, lo, hi)
      pending.append(dist.all_reduce(out_full[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
    for work in pending:
      work.wait()
    return out_full


class _DeviceView:
  """Lets torch wrap device memory owned by the C library (no copy)."""

  def __init__(self, ptr, n):
    self.__cuda_array_interface__ = dict(shape=(int(n),), typestr='<f8', data=(int(ptr), False), version=3)


class PeerExchangeRhs:
  """dy/dt of the full problem on every rank with the flux exchange done by this library's own
  kernels over NVLink peer memory (ranks = GPUs of one node, one process each).

  States are owned in contiguous blocks.  The product kernel stores this rank's partial dy/dt of a
  state straight into the owner's staging buffer while it computes; in `rounds` rounds the owners
  add their `world` slots in rank order and store the sums into every rank's result vector, one
  round behind the product (tapes_peer_rhs).  Cross-GPU ordering uses epoch flags in peer memory,
  so no NCCL kernel competes with the memory-bound product for SMs; torch.distributed only
  carries the 64-byte CUDA IPC handles at set-up.  Every rank ends up with bit-identical dy/dt.
  """

  def __init__(self, model, group=None, rounds=4):
    import ctypes
    from . import _lib, markov_tapes
    self.lib = markov_tapes.u_lib
    self.model = model
    self.world = dist.get_world_size(group)
    self.rank = dist.get_rank(group)
    self.n = model.n_states
    self.rounds = max(1, min(16, int(rounds)))
    unit = 32 * self.rounds
    block = -(-self.n // self.world)
    self.block = -(-block // unit) * unit
    self.padded = self.block * self.world
    sizes = (self.padded, self.padded, 2 * self.world)  # staging, result, flags
    handles = [(ctypes.c_ubyte * 64)() for _ in sizes]
    self._own, self._opened, self.group, self.out = [], [], None, None
    # every step below ends in a collective that all ranks reach whether or not their own part
    # worked, so a rank that cannot allocate or map peer memory makes all ranks raise together
    # instead of leaving the others waiting
    problem = None
    try:
      self._own = [self.lib.tapes_peer_alloc(n, h) for n, h in zip(sizes, handles)]
      _lib.check(all(bool(p) for p in self._own), 'tapes_peer_alloc')
    except Exception as ex:  # pylint: disable=broad-except
      problem = repr(ex)
    everyone = [None] * self.world
    dist.all_gather_object(everyone, (problem, tuple(bytes(h) for h in handles)), group=group)
    self._raise_together([e[0] for e in everyone], group, collective_done=True)
    try:
      tables = [[], [], []]
      for r, (_, theirs) in enumerate(everyone):
        for kind, handle in enumerate(theirs):
          if r == self.rank:
            tables[kind].append(self._own[kind])
          else:
            ptr = self.lib.tapes_peer_open(handle)  # bytes: ctypes passes the address of the 64-byte buffer
            _lib.check(bool(ptr), 'tapes_peer_open')
            self._opened.append(ptr)
            tables[kind].append(ptr)
      arrays = [(ctypes.c_void_p * self.world)(*t) for t in tables]
      self.group = self.lib.tapes_peer_group_create(self.world, self.rank, self.block, self.rounds, *arrays)
      _lib.check(bool(self.group), 'tapes_peer_group_create')
      device = torch.device('cuda', torch.cuda.current_device())
      self.out = torch.as_tensor(_DeviceView(self._own[1], self.padded), device=device)
    except Exception as ex:  # pylint: disable=broad-except
      problem = repr(ex)
    # also the barrier "every rank has mapped every buffer" before anyone stores into them
    self._raise_together([problem], group, collective_done=False)

  def _raise_together(self, problems, group, collective_done):
    if not collective_done:
      gathered = [None] * self.world
      dist.all_gather_object(gathered, problems[0], group=group)
      problems = gathered
    bad = [(r, p) for r, p in enumerate(problems) if p]
    if bad:
      self._release()
      raise RuntimeError('peer exchange set-up failed on rank(s) ' + '; '.join(f'{r}: {p}' for r, p in bad))

  def _release(self):
    if self.group:
      self.lib.tapes_peer_group_destroy(self.group)
    self.group = None
    for p in self._opened:
      self.lib.tapes_peer_close(p)
    self._opened = []
    self.out = None
    for p in self._own:
      if p:
        self.lib.tapes_peer_free(p)
    self._own = []

  def rhs_full(self, p_full):
    """p_full: at least n_states doubles on this device.  Returns the summed dy/dt (a view of the
    peer-visible result vector, valid until the next call).  Collective: every rank calls it."""
    from . import _lib, device as dev
    rc = self.lib.tapes_peer_rhs(self.group, self.model.handle, p_full.data_ptr(), dev._current_stream_handle())
    _lib.check(rc == 0, 'tapes_peer_rhs')
    return self.out

  def check(self):
    """Raises when a cross-GPU wait timed out (results after that point are not valid)."""
    torch.cuda.synchronize()
    if self.lib.tapes_peer_group_error(self.group):
      raise RuntimeError('peer exchange: a rank did not signal within the timeout')

  def close(self, group=None):
    torch.cuda.synchronize()
    dist.barrier(group=group)  # nobody is still storing into a buffer that is about to go away
    self._release()
    dist.barrier(group=group)


class OverlappedRhs:
  """dy/dt of the full problem with the exchange overlapped chunk by chunk.

  local_weights(p_full) evaluates everything that depends on p; local_flux_rows(out, lo, hi)
  writes this rank's partial dy/dt for the states lo <= i < hi.  `owner_update(block, lo, hi)`
  (optional) is applied by the owner to its summed block before it is gathered - the place where
  an integrator turns flux into a new table; the default leaves dy/dt in place.
  """

  def __init__(self, local_weights, local_flux_rows, n_states, chunks=8, group=None, device='cuda',
               owner_update=None):
    self.local_weights = local_weights
    self.local_flux_rows = local_flux_rows
    self.owner_update = owner_update
    self.group = group
    self.world = dist.get_world_size(group)
    self.rank = dist.get_rank(group)
    self.n = n_states
    self.chunks = max(1, int(chunks))
    unit = self.chunks * self.world
    self.padded = -(-n_states // unit) * unit
    self.chunk_len = self.padded // self.chunks
    self.block_len = self.chunk_len // self.world
    self.partial = torch.zeros(self.padded, dtype=torch.float64, device=device)
    self.owned = torch.zeros(self.chunks, self.block_len, dtype=torch.float64, device=device)

  def owned_range(self, chunk):
    lo = chunk * self.chunk_len + self.rank * self.block_len
    return lo, lo + self.block_len

  def rhs_full(self, p_full, out_full):
    """p_full, out_full: padded full-length vectors (padding stays zero)."""
    self.local_weights(p_full[:self.n])
    pending = []
    for c in range(self.chunks):
      lo = c * self.chunk_len
      hi = min(lo + self.chunk_len, self.n)
      if hi > lo:
        self.local_flux_rows(self.partial, lo, hi)
      seg = self.partial[c * self.chunk_len:(c + 1) * self.chunk_len]
      w1 = dist.reduce_scatter_tensor(self.owned[c], seg, op=dist.ReduceOp.SUM, group=self.group,
                                      async_op=True)
      pending.append((c, w1))
      if len(pending) > 1:
        self._finish(pending.pop(0), out_full)
    while pending:
      self._finish(pending.pop(0), out_full)
    for w in self._gathers:
      w.wait()
    self._gathers = []
    return out_full

  _gathers = []

  def _finish(self, item, out_full):
    c, w1 = item
    w1.wait()
    if self.owner_update is not None:
      self.owner_update(self.owned[c], *self.owned_range(c))
    seg = out_full[c * self.chunk_len:(c + 1) * self.chunk_len]
    self._gathers = self._gathers + [dist.all_gather_into_tensor(seg, self.owned[c], group=self.group,
                                                                 async_op=True)]
