"""Multi-GPU sharding of the master-equation step (one process per GPU, torch.distributed).

The window-extension forests of different flux rules are independent, so the rule set is dealt
round-robin to the ranks: rank g builds and evaluates only its rules over the full state space
and produces a partial dy/dt.  States are owned in contiguous blocks: a reduce-scatter sums the
partial flux into each owner's block (the flux exchange), the owner applies its update, and an
all-gather rebuilds the full table for the next right-hand side.  Nothing here touches the GPU
directly, so the plumbing is testable with the gloo backend on CPU.
"""

import numpy
import torch
import torch.distributed as dist


def split_rule_set(rules, world_size, rank):
  """Rules rank, rank + world_size, ... of a rule-set dict, plus one inert rule that carries the
  selection weight of everything dealt to other ranks, so that the probability of picking a local
  rule is the same as in the unsplit problem (gambit_macros.scm:75-86 normalises by the sum)."""
  n = len(rules['rate'])
  mine = numpy.arange(rank, n, world_size)
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
