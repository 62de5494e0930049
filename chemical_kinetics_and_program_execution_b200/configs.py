"""Workload definitions: the reference examples' initial distributions (vectorised restatements
of the `get_p0` generators in /examples, checked against fixtures produced by the reference's own
generators in tests/golden/), product tables, and the synthetic random rewrite-rule sets of
BASELINE.json's last config.  Imports without a GPU.
"""

import numpy


def _flat_index(digits, size_a):
  idx = 0
  for d in digits:
    idx = idx * size_a + d
  return idx


def _one_foreign_symbol(size_a, cl_k, background, symbol):
  """Flat indices of the cl_k sequences that are `background` everywhere except one `symbol`."""
  base = _flat_index([background] * cl_k, size_a)
  return numpy.array([base + (symbol - background) * size_a ** (cl_k - 1 - j)
                      for j in range(cl_k)], dtype=numpy.int64)


def ex2_p0(cl_k, p_pair=0.01):
  """examples/ex2_ferromagnet_tape.py:43-52."""
  p0 = numpy.zeros(2 ** cl_k, dtype=numpy.float64)
  p0[0] = 1.0 - p_pair * (cl_k + 1)
  for j in range(cl_k - 1):
    p0[0b11 << j] = p_pair
  p0[1] = p_pair
  p0[1 << (cl_k - 1)] = p_pair
  return p0


def ex3_p0(cl_k=6, p_a=0.02):
  """examples/ex3_copolymerization.py:38-53 (alphabet O A M N)."""
  sym_o, sym_a, sym_m, sym_n = range(4)
  p0 = numpy.zeros(4 ** cl_k, dtype=numpy.float64)
  p0[0] = 1 - cl_k * p_a * 2
  p0[_one_foreign_symbol(4, cl_k, sym_o, sym_a)] = p_a
  p0[_one_foreign_symbol(4, cl_k, sym_o, sym_m)] = 0.5 * p_a
  p0[_one_foreign_symbol(4, cl_k, sym_o, sym_n)] = 0.5 * p_a
  return p0


def ex4_p0(cl_k=5, tape_fraction=0.25, cursor_fraction=0.01, powered_fraction=0.05):
  """examples/ex4_chemical_turing.py:44-83 with random01=False (alphabet A B C D I O P X S)."""
  size_a = 9
  sym_a, sym_o, sym_p, sym_s = 0, 5, 6, 8
  p0 = numpy.zeros(size_a ** cl_k, dtype=numpy.float64)
  p0[_one_foreign_symbol(size_a, cl_k, sym_s, sym_p)] = (1 - tape_fraction) * powered_fraction
  p0[_flat_index([sym_s] * cl_k, size_a)] = (1 - tape_fraction) * (1 - powered_fraction * cl_k)
  p0[_one_foreign_symbol(size_a, cl_k, sym_o, sym_a)] = tape_fraction * cursor_fraction
  p0[_flat_index([sym_o] * cl_k, size_a)] = tape_fraction * (1 - cursor_fraction * cl_k)
  return p0


def ex4var2_p0(cl_k=5, tape_fraction=0.25, cursor_fraction=0.04, powered_fraction=0.1):
  """examples/ex4var2_chemical_turing.py:88-115 `get_p0e`, random01=False (adds symbol E)."""
  size_a = 10
  sym_o, sym_p, sym_s, sym_e = 5, 6, 8, 9
  p0 = numpy.zeros(size_a ** cl_k, dtype=numpy.float64)
  p0[_flat_index([sym_s] * cl_k, size_a)] = (1 - tape_fraction) * (
      1 - powered_fraction * cl_k - cursor_fraction * cl_k)
  p0[_one_foreign_symbol(size_a, cl_k, sym_s, sym_p)] = (1 - tape_fraction) * powered_fraction
  p0[_one_foreign_symbol(size_a, cl_k, sym_s, sym_e)] = (1 - tape_fraction) * cursor_fraction
  p0[_flat_index([sym_o] * cl_k, size_a)] = tape_fraction
  return p0


def ex5_p0(cl_k=5):
  """examples/ex5_msrtf_machine.py:45-49: uniform over the first three of five symbols."""
  p0 = numpy.zeros([5] * cl_k, dtype=numpy.float64)
  p0[(slice(0, 3),) * cl_k] = 3.0 ** (-cl_k)
  return p0.ravel()


def product_table(freqs, cl_k):
  """Subsequence table of an i.i.d. tape with symbol frequencies `freqs` (shift consistent)."""
  freqs = numpy.asarray(freqs, dtype=numpy.float64)
  table = freqs
  for _ in range(cl_k - 1):
    table = numpy.multiply.outer(table, freqs)
  return table.ravel()


def dirichlet_product_table(size_a, cl_k, seed):
  """Full-support product table with Dirichlet(1) symbol frequencies."""
  rng = numpy.random.default_rng(seed)
  return product_table(rng.dirichlet(numpy.ones(size_a)), cl_k)


def markov_table(size_a, cl_k, seed):
  """Subsequence table of a random first-order Markov chain in its stationary state: shift
  consistent but not a product table."""
  rng = numpy.random.default_rng(seed)
  trans = rng.dirichlet(numpy.ones(size_a), size=size_a)  # trans[a, b] = P(b | a)
  evals, evecs = numpy.linalg.eig(trans.T)
  stat = numpy.real(evecs[:, numpy.argmin(abs(evals - 1))])
  stat = stat / stat.sum()
  table = stat
  for _ in range(cl_k - 1):
    table = table[..., :, None] * trans.reshape((1,) * (table.ndim - 1) + trans.shape)
  return table.ravel()


def random_rule_set(size_a, n_rules, seed, max_span=3, catalyst_fraction=0.5):
  """Synthetic random rewrite rules.

  Rule r rewrites `span` cells of one tape from `pattern` to `repl` (different in at least one
  cell) with acceptance rate u_r in (0, 1]; with probability `catalyst_fraction` it additionally
  requires a given symbol under the head of the other tape, which makes its rate depend on p.
  """
  rng = numpy.random.default_rng(seed)
  tape = rng.integers(0, 2, size=n_rules)
  span = rng.integers(1, max_span + 1, size=n_rules)
  catalyst = numpy.where(rng.random(n_rules) < catalyst_fraction,
                         rng.integers(0, size_a, size=n_rules), -1)
  pattern = numpy.zeros((n_rules, 4), dtype=numpy.int32)
  repl = numpy.zeros((n_rules, 4), dtype=numpy.int32)
  for r in range(n_rules):
    m = span[r]
    pattern[r, :m] = rng.integers(0, size_a, size=m)
    while True:
      repl[r, :m] = rng.integers(0, size_a, size=m)
      if (repl[r, :m] != pattern[r, :m]).any():
        break
  rate = 1.0 - rng.random(n_rules)  # (0, 1]
  return dict(tape=tape.astype(numpy.int32), span=span.astype(numpy.int32),
              catalyst=catalyst.astype(numpy.int32), pattern=pattern, repl=repl,
              rate=rate, select_weight=numpy.ones(n_rules))


def rotated_rule_set(rules, shift, size_a):
  """The same rules with every symbol s replaced by (s + shift) mod size_a: a different rule set
  with exactly the same structure sizes (used to give every GPU of a weak-scaling run equal work)."""
  out = {key: numpy.array(val, copy=True) for key, val in rules.items()}
  span = numpy.asarray(rules['span'])
  for r in range(len(span)):
    m = int(span[r])
    out['pattern'][r, :m] = (out['pattern'][r, :m] + shift) % size_a
    out['repl'][r, :m] = (out['repl'][r, :m] + shift) % size_a
    if out['catalyst'][r] >= 0:
      out['catalyst'][r] = (out['catalyst'][r] + shift) % size_a
  return out


def concat_rule_sets(parts):
  return {key: numpy.concatenate([numpy.asarray(p[key]) for p in parts]) for key in parts[0]}


def autocatalysis_rule_set(c_form=0.01, c_auto=1.0, c_stab=0.05, c_add=0.02, c_remove=0.02):
  """A tape restatement of the chemistry of examples/autocatalysis.py (BASELINE.json config 2).

  The reference script integrates a 3-variable mean-field ODE and never touches the tape path
  (examples/autocatalysis.py:38-44, 126-151), so this rule set is OUR definition, not a port:
  alphabet (S M A B) = solvent, monomer, A-dimer half, B-dimer half.  Two adjacent monomers on the
  data tape dimerise spontaneously (c_form) or catalysed by a dimer half of the same kind under
  the program-tape head (c_auto); dimers fall apart (c_stab); monomers flow in and out
  (c_add / c_remove).  Parity for it is against the CPU oracle only.
  """
  S, M, A, B = range(4)
  rows = []  # tape, span, catalyst, pattern, repl, rate
  for dimer in (A, B):
    rows.append((1, 2, -1, [M, M], [dimer, dimer], c_form))
    rows.append((1, 2, dimer, [M, M], [dimer, dimer], c_auto))
    rows.append((1, 2, -1, [dimer, dimer], [M, M], c_stab))
  rows.append((1, 1, -1, [S], [M], c_add))
  rows.append((1, 1, -1, [M], [S], c_remove))
  n = len(rows)
  pattern = numpy.zeros((n, 4), dtype=numpy.int32)
  repl = numpy.zeros((n, 4), dtype=numpy.int32)
  for r, row in enumerate(rows):
    pattern[r, :row[1]] = row[3]
    repl[r, :row[1]] = row[4]
  return dict(tape=numpy.array([r[0] for r in rows], dtype=numpy.int32),
              span=numpy.array([r[1] for r in rows], dtype=numpy.int32),
              catalyst=numpy.array([r[2] for r in rows], dtype=numpy.int32),
              pattern=pattern, repl=repl, rate=numpy.array([r[5] for r in rows], dtype=numpy.float64),
              select_weight=numpy.ones(n))


def synthetic_tag(size_a, n_rules, seed, max_span=3):
  return f'synthetic-A{size_a}-R{n_rules}-s{seed}-m{max_span}'


# The reference's example configurations (tag, alphabet, cl_k, stepper settings).
EXAMPLES = {
    'ex2': dict(tag='ex2-ferromagnetic-chain', size_a=2, cl_k=(3, 4, 5, 6, 7), t_max=60.0,
                n_ts=1001, method='odeint', rtol=1e-9, atol=1e-9, p_pair=1 / 250),
    'ex3': dict(tag='ex3-copolymerization', size_a=4, cl_k=6, t_max=1000.0, n_ts=1001,
                method='odeint', rtol=1e-9, atol=1e-9),
    'ex4': dict(tag='ex4-chemical-turing', size_a=9, cl_k=5, t_max=2000.0, n_ts=2001,
                method='DOP853', rtol=1e-13, atol=1e-13),
    'ex5': dict(tag='ex5-msrtf-machine', size_a=5, cl_k=5, t_max=500.0, n_ts=4001,
                method='DOP853', rtol=1e-13, atol=1e-13),
}
