"""NumPy restatement of the Monte-Carlo simulator's rules (csrc/montecarlo.cu of the product).

TEST INFRASTRUCTURE ONLY, like the rest of oracle/: the simulator is this repo's own addition
(SURVEY.md section 8(f) rank 4), so its checker is a second implementation of the same rules -
counter-based random numbers, simultaneous events per sub-step, highest event number wins a
contested cell - against which the CUDA kernels are compared bit for bit, while the physics is
checked against the master equation integrated through the oracle.
"""

import numpy
U64 = numpy.uint64
M1, M2, GOLD = U64(0xff51afd7ed558ccd), U64(0xc4ceb9fe1a85ec53), U64(0x9E3779B97F4A7C15)

def mix(x):
  x = x ^ (x >> U64(33)); x = x * M1
  x = x ^ (x >> U64(33)); x = x * M2
  return x ^ (x >> U64(33))

def draw(seed, substep, event, d):
  with numpy.errstate(over='ignore'):
    x = mix(U64(seed) + GOLD)
    x = mix(x ^ U64(substep))
    x = mix(x + numpy.asarray(event, dtype=U64) * GOLD)
    return mix(x ^ U64(d))

def substep(tree, size_a, tape, n_events, seed, step):
  N = tape.size
  e = numpy.arange(n_events, dtype=U64)
  heads = numpy.stack([(draw(seed, step, e, 0) % U64(N)).astype(numpy.int64),
                       (draw(seed, step, e, 1) % U64(N)).astype(numpy.int64)])  # program head, data head
  node = numpy.zeros(n_events, dtype=numpy.int64)
  W = 8
  wpos = numpy.full((n_events, W), -1, dtype=numpy.int64); wsym = numpy.zeros((n_events, W), dtype=numpy.int64)
  wcnt = numpy.zeros(n_events, dtype=numpy.int64); picks = numpy.zeros(n_events, dtype=numpy.int64)
  kind, a, b, c = tree['kind'], tree['a'], tree['b'], tree['c']
  fc, fw, child, weight = tree['first_child'], tree['first_weight'], tree['child'], tree['weight']
  ev = numpy.arange(n_events)
  while True:
    k = kind[node]
    if (k == 0).all(): break
    rd = k == 1
    if rd.any():
      idx = ev[rd]; nd = node[idx]
      pos = (heads[a[nd], idx] + b[nd]) % N
      sym = tape[pos].astype(numpy.int64)
      for w in range(W):  # own writes, later ones override earlier ones
        hit = (wpos[idx, w] == pos) & (w < wcnt[idx])
        sym = numpy.where(hit, wsym[idx, w], sym)
      node[idx] = child[fc[nd] + sym]
    wr = k == 2
    if wr.any():
      idx = ev[wr]; nd = node[idx]
      pos = (heads[a[nd], idx] + b[nd]) % N
      assert (wcnt[idx] < W).all()
      wpos[idx, wcnt[idx]] = pos; wsym[idx, wcnt[idx]] = c[nd]; wcnt[idx] += 1
      node[idx] = child[fc[nd]]
    pk = k == 3
    if pk.any():
      idx = ev[pk]; nd = node[idx]
      u = (draw(seed, step, idx.astype(U64), 2 + picks[idx]) >> U64(11)).astype(numpy.float64) * 2.0 ** -53
      picks[idx] += 1
      choice = numpy.zeros(idx.size, dtype=numpy.int64)
      for n, (i_, nd_, u_) in enumerate(zip(idx, nd, u)):
        wts = weight[fw[nd_]:fw[nd_] + a[nd_]]
        total = 0.0
        for x in wts: total = total + x
        x = u_ * total; cum = 0.0; j = a[nd_] - 1
        for jj, wj in enumerate(wts):
          cum = cum + wj
          if x < cum: j = jj; break
        choice[n] = j
      node[idx] = child[fc[nd] + choice]
  stamp = numpy.zeros(N, dtype=numpy.int64)
  for w in range(W):
    live = w < wcnt
    val = ((ev[live] + 1) << 12) | (w << 8) | wsym[live, w]
    numpy.maximum.at(stamp, wpos[live, w], val)
  out = tape.copy()
  hit = stamp > 0
  out[hit] = (stamp[hit] & 255).astype(tape.dtype)
  return out

def window_counts(tape, size_a, k):
  N = tape.size
  idx = numpy.zeros(N, dtype=numpy.int64)
  for o in range(k):
    idx = idx * size_a + numpy.roll(tape, -o).astype(numpy.int64)
  return numpy.bincount(idx, minlength=size_a ** k)

def sample_ring(p, size_a, k, N, rng):
  p = numpy.asarray(p).reshape([size_a] * k)
  tape = numpy.zeros(N, dtype=numpy.uint8)
  first = rng.choice(size_a ** k, p=p.ravel() / p.sum())
  digits = numpy.unravel_index(first, [size_a] * k)
  tape[:k] = digits
  cond = p / numpy.maximum(p.sum(axis=-1, keepdims=True), 1e-300)
  u = rng.random(N)
  for n in range(k, N):
    row = cond[tuple(tape[n - k + 1:n])]
    tape[n] = min(numpy.searchsorted(numpy.cumsum(row), u[n]), size_a - 1)
  return tape


def ferromagnet_chain_counts(chain0, sites, uniforms, accept):
  """NumPy restatement of the reference's ferromagnet Monte Carlo for given random numbers
  (examples/ex2_ferromagnet_mc.py:46-122 `simulate`, 134-163 `island_length_stats`, 182-189): the
  checker of tapes_mc_ferromagnet_chains.  chain0 [T, N], sites / uniforms [T, S-1, M], accept [3, 2];
  returns counts [T, S, 6].  The trials of a time step all read the state of the previous step and
  xor into the new one (ex2_ferromagnet_mc.py:95-121: `result[nt-1, stride3]`, `result[nt, ...] ^= 1`)."""
  chain0 = numpy.asarray(chain0).astype(numpy.int64)
  n_trials, n = chain0.shape
  n_steps = sites.shape[1] + 1
  counts = numpy.zeros((n_trials, n_steps, 6))
  for t in range(n_trials):
    cur = chain0[t].copy()
    for step in range(n_steps):
      if step > 0:
        prev = cur.copy()
        i = numpy.asarray(sites[t, step - 1], dtype=numpy.int64)
        mid = prev[i]
        equal = (prev[(i - 1) % n] == mid).astype(numpy.int64) + (prev[(i + 1) % n] == mid)
        flips = uniforms[t, step - 1] < accept[equal, mid]
        numpy.bitwise_xor.at(cur, i[flips], 1)
      starts = numpy.nonzero((cur == 1) & (numpy.roll(cur, 1) == 0))[0]
      length = numpy.ones(starts.size, dtype=numpy.int64)
      alive = numpy.ones(starts.size, dtype=bool)
      for d in range(1, 6):
        alive &= cur[(starts + d) % n] == 1
        length += alive
      for c_len in range(1, 6):
        counts[t, step, c_len] = numpy.count_nonzero(length == c_len)
  return counts
