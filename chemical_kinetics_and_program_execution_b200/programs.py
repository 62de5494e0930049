"""Tape programs written as Python functions.

The reference states a problem as a Scheme body over three primitives - `tape-get`, `tape-set!`,
`choose` (framework/gambit_macros.scm:99-125; bodies in framework/problems.scm) - and needs
problems.scm edited and the shared object rebuilt for every new problem (MAKE.sh:43-47).  Here a
body is a Python function of one argument offering the same three primitives; `trace` runs it
against every combination of outcomes and records its decision tree, which the library takes as
data (tapes_register_program, include/tapes_b200.h).  The body must be deterministic apart from
what it reads and chooses, like the reference's bodies, which are re-entered through continuations
(framework/tape_multiverse.scm:750-838).

  def decay(tape):                       # framework/problems.scm:22-26
    if tape.get(True, 0) == 1:
      tape.set(True, 0, 0)
  markov_tapes.register_program('my-decay', 2, decay)
"""

import numpy

END, READ, WRITE, PICK = 0, 1, 2, 3


class _Fork(Exception):
  def __init__(self, ways):
    super().__init__(ways)
    self.ways = ways


class Tape:
  """What a body sees.  data_tape: False = program tape, True = data tape (the reference's
  data-tape? flag); cell indices are relative to the head."""

  def __init__(self, size_a, decisions):
    self._size_a = size_a
    self._decisions = decisions
    self._used = 0
    self._known = ({}, {})
    self.ops = []

  def _decide(self, ways):
    if self._used >= len(self._decisions):
      raise _Fork(ways)
    self._used += 1
    return self._decisions[self._used - 1]

  def get(self, data_tape, index):
    """The symbol (an alphabet index) in cell `index`; a body sees its own writes
    (framework/tape_multiverse.scm:759)."""
    t, index = (1 if data_tape else 0), int(index)
    if index not in self._known[t]:
      self.ops.append((READ, t, index))
      self._known[t][index] = self._decide(self._size_a)
    return self._known[t][index]

  def set(self, data_tape, index, symbol):
    t, index, symbol = (1 if data_tape else 0), int(index), int(symbol)
    if not 0 <= symbol < self._size_a:
      raise ValueError(f'symbol {symbol} outside the alphabet of {self._size_a}')
    self.ops.append((WRITE, t, index, symbol))
    self._known[t][index] = symbol

  def choose(self, weights):
    """Index of the option picked with probability weights[j] / sum(weights)
    (framework/gambit_macros.scm:75-86, 119-124)."""
    weights = tuple(float(w) for w in weights)
    if not weights:
      raise ValueError('choose needs at least one option')
    self.ops.append((PICK,) + weights)
    return self._decide(len(weights))

  def choose_value(self, options):
    """options: (weight, value) pairs like the reference's (choose '((1.0 #t) (1.0 #f)))."""
    options = list(options)
    return options[self.choose([w for w, _ in options])][1]


def trace(body, size_a, max_nodes=1 << 22):
  """Decision tree of `body` as the dict of arrays tapes_register_program takes."""
  root = {}
  todo = [[]]
  n_nodes = 0
  while todo:
    decisions = todo.pop()
    tape = Tape(size_a, decisions)
    try:
      body(tape)
    except _Fork as fork:
      todo.extend(decisions + [c] for c in range(fork.ways - 1, -1, -1))
      continue
    # a complete run: thread its operations into the tree
    node, used = root, 0
    for op in tape.ops:
      if 'op' not in node:
        fan = size_a if op[0] == READ else (1 if op[0] == WRITE else len(op) - 1)
        node.update(op=op, children=[{} for _ in range(fan)])
        n_nodes += 1
        if n_nodes > max_nodes:
          raise ValueError('program tree too large')
      elif node['op'] != op:
        raise ValueError('the body is not deterministic: the same outcomes led to different operations')
      if op[0] == WRITE:
        node = node['children'][0]
      else:
        node = node['children'][decisions[used]]
        used += 1
    if 'op' in node:
      raise ValueError('the body is not deterministic: the same outcomes led to different operations')
    node['end'] = True
  # flatten in preorder; every run ends in the shared END node, which comes last
  order, stack = [], [root]
  while stack:
    node = stack.pop()
    if 'op' in node:
      node['id'] = len(order)
      order.append(node)
      stack.extend(reversed(node['children']))
  end_id = len(order)
  n = end_id + 1
  kind = numpy.zeros(n, dtype=numpy.int32)
  a, b, c = (numpy.zeros(n, dtype=numpy.int32) for _ in range(3))
  first_child = numpy.zeros(n, dtype=numpy.int32)
  first_weight = numpy.zeros(n, dtype=numpy.int32)
  child, weight = [], []
  for node in order:
    i, op = node['id'], node['op']
    kind[i] = op[0]
    first_child[i] = len(child)
    child.extend(ch['id'] if 'op' in ch else end_id for ch in node['children'])
    if op[0] == PICK:
      a[i] = len(op) - 1
      first_weight[i] = len(weight)
      weight.extend(op[1:])
    else:
      a[i], b[i] = op[1], op[2]
      if op[0] == WRITE:
        c[i] = op[3]
  return dict(kind=kind, a=a, b=b, c=c, first_child=first_child, first_weight=first_weight,
              child=numpy.array(child, dtype=numpy.int32), weight=numpy.array(weight, dtype=numpy.float64))
