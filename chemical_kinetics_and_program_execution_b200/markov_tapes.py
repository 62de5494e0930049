"""Drop-in replacement for the reference's framework/markov_tapes.py, backed by CUDA on a B200.

Same public names, keyword-only signatures, return shapes and error behaviour as the reference
module (framework/markov_tapes.py:58-374); the Gambit-C shared object behind it is replaced by
`tapes_py_interface.so` built from csrc/ (C ABI: include/tapes_b200.h).  Like the reference,
importing this module initialises the runtime, registers the problems and runs the load-time
known-answer test (framework/markov_tapes.py:357-374), so it needs a GPU and raises otherwise.

Differences, all documented in INTEGRATION.md:
  * the per-call `print('DDD t=...')` (framework/markov_tapes.py:277) only happens when
    MARKOV_TAPES_DEBUG=1;
  * a failed library call raises RuntimeError (the reference drops into a Scheme REPL,
    framework/tapes_py_interface.scm:42-44);
  * additions: `ode_integrate_device`, `sequence_observable`, `model_stats`, `check_table`,
    `register_rule_set`, `register_program`, `monte_carlo`, `ferromagnet_monte_carlo`; device-resident
    dy/dt lives in device.py.
"""

import atexit
import itertools
import os
import sys
import types

import numpy
import scipy.integrate

from . import _lib

IS_DEBUG = bool(int(os.getenv('MARKOV_TAPES_DEBUG', '0')))

u_lib = _lib.load()


def init_gambit():
  """Initialises the device runtime (name kept from the reference, markov_tapes.py:58-76)."""
  if IS_DEBUG:
    print('DEBUG: initializing runtime.', file=sys.stderr, flush=True)
  runtime = u_lib.setup_gambit()
  if not runtime:
    _lib.check(False, 'setup_gambit')

  def _cleanup_gambit():
    if IS_DEBUG:
      print('DEBUG: cleanup runtime.', file=sys.stderr, flush=True)
    u_lib.cleanup_gambit(runtime)
  atexit.register(_cleanup_gambit)
  if IS_DEBUG:
    print('DEBUG: registering problems.', file=sys.stderr, flush=True)
  if 124 != u_lib.c_register_problems(123):
    # canary: the call returns its argument plus one (markov_tapes.py:72-76)
    raise ValueError('Registering problems failed.')


### Helpers (host-side analysis; same semantics as markov_tapes.py:81-256)

def mpp_from_spd(spd, eps=None):
  """Markov process parameters r[prefix + (s,)] = P(s | prefix) from a subsequence table.

  spd has shape B + (N,)*k.  Entries are clipped to [eps, 1] first so that an impossible prefix
  yields a uniform continuation instead of 0/0 (default eps 1e-100).
  """
  if eps is None:
    eps = 1e-100
  clipped = numpy.clip(numpy.asarray(spd).astype(numpy.float64), eps, 1)
  return clipped / clipped.sum(axis=-1, keepdims=True)


def ctm_from_mpp(num_alphabet, num_context, mpp):
  """Context transfer matrix [N**c, N**c]: entry [suffix context, prefix context]."""
  n_ctx = num_alphabet ** num_context
  steps = numpy.asarray(mpp).reshape([num_alphabet] * (1 + num_context))
  # column = context (i_0..i_{c-1}), row = shifted context (i_1..i_c)
  cols = numpy.arange(n_ctx * num_alphabet) // num_alphabet
  rows = numpy.arange(n_ctx * num_alphabet) % n_ctx
  result = numpy.zeros([n_ctx, n_ctx])
  numpy.add.at(result, (rows, cols), steps.ravel())
  return result


def get_ctm_eigenvalue1_eigenspace(spd, eps_mpp=None, eps=1e-7):
  """Eigenvalue-1 eigenspace of the context transfer matrix (markov_tapes.py:133-175).

  Returns (deviation, eigenspace) or (marginals_distance, None) when the marginals over the
  leading and the trailing index differ by more than eps.
  """
  spd = numpy.asarray(spd, dtype=numpy.float64)
  num_alphabet = spd.shape[0]
  num_context = spd.ndim - 1
  right = spd.sum(axis=-1)
  left = spd.sum(axis=0)
  distance = numpy.linalg.norm(right.ravel() - left.ravel())
  if not distance <= eps:
    return distance, None
  ctm = ctm_from_mpp(num_alphabet, num_context, mpp_from_spd(spd, eps=eps_mpp))
  eigvals, eigvecs = numpy.linalg.eig(ctm)
  eigenspace = eigvecs[:, abs(eigvals - 1.0) <= eps]
  _, residuals, *_ = numpy.linalg.lstsq(eigenspace, left.ravel(), rcond=None)
  return numpy.linalg.norm(residuals**.5), eigenspace


def markov_entropy(spd):
  """Entropy rate of the Markov chain described by the subsequence table."""
  clipped = numpy.clip(numpy.asarray(spd).astype(numpy.float64), 1e-280, 1)
  context = clipped.sum(axis=-1)
  conditional = clipped / context[..., numpy.newaxis]
  per_context = (-conditional * numpy.log(conditional)).sum(axis=-1)
  return per_context.ravel().dot(context.ravel())


def seq_prob(spd, seq, *, num_prefix_indices=0, eps=None, mpp=None, want_mpp=False):
  """Probability of `seq` under the subsequence table (markov_tapes.py:190-233).

  Returns (probability, mpp).  Sequences not longer than k are read off directly (summing the
  leading excess axes); longer ones are extended with the Markov process parameters.
  """
  spd = numpy.asarray(spd, dtype=numpy.float64)
  k = spd.ndim - num_prefix_indices
  excess = k - len(seq)
  if excess >= 0:
    picked = spd[(Ellipsis,) + tuple(seq)]
    axes = tuple(range(num_prefix_indices, num_prefix_indices + excess))
    return picked.sum(axis=axes), (mpp_from_spd(spd, eps=eps) if want_mpp else mpp)
  if mpp is None:
    mpp = mpp_from_spd(spd, eps=eps)
  current = spd[(Ellipsis,) + tuple(seq[:k])]
  rest = seq[1:]
  while len(rest) >= k:
    current = mpp[(Ellipsis,) + tuple(rest[:k])] * current
    rest = rest[1:]
  return current, mpp


def tprint(size_a, cl_k, adata, epsilon=1e-10, nmax=float('inf'), file=None):
  """Prints the entries of a transition table whose magnitude is not below epsilon."""
  n_in = cl_k - 1
  table = numpy.asarray(adata).reshape([size_a] * (2 * n_in))
  for n, idx in enumerate(itertools.product(range(size_a), repeat=2 * n_in)):
    if n >= nmax:
      print('... more entries...', file=file)
    val = table[idx]
    if not abs(val) < epsilon:
      print(f'{idx[:n_in]} {idx[n_in:]}: {val}')


### Right-hand side and integrators

_PINNED_RESULT_STATES = 1 << 22


def _tag_buffer(tag):
  return numpy.frombuffer(tag.encode() + b'\x00', dtype=numpy.uint8)


def get_dy_dt(*, tag, size_a, cl_k, debug=False):
  """Returns the (probabilities_in, t) -> d/dt probabilities function (markov_tapes.py:259-289).

  Host arrays in, host array out; the work happens on the GPU through c_compute_dy_dt.
  """
  a_tag = _tag_buffer(tag)
  do_debug = 1 if debug else 0
  expected_size = size_a ** cl_k
  # large tables: results in page-locked arrays (plain DMA, no zero-filling of 8 * n fresh bytes per
  # call); the library overwrites every entry, or fills the array with NaN when it fails
  pinned = _lib.PinnedResults(expected_size) if expected_size >= _PINNED_RESULT_STATES else None

  def dy_dt(a_probs_in, t):
    if IS_DEBUG:
      print(f'DDD {t=:.10g}')
    c_probs_in = numpy.ascontiguousarray(numpy.asarray(a_probs_in, dtype=numpy.float64).ravel())
    if c_probs_in.size != expected_size:
      raise ValueError(f'probability-array should have size {expected_size}, '
                       f'observed: {c_probs_in.size}')
    c_probs_out = pinned.take() if pinned is not None else None
    if c_probs_out is None:
      c_probs_out = numpy.zeros_like(c_probs_in)
    u_lib.c_compute_dy_dt(a_tag.ctypes.data, cl_k, do_debug, c_probs_in.ctypes.data,
                          c_probs_out.ctypes.data)
    if u_lib.tapes_last_error():
      _lib.check(False, 'c_compute_dy_dt')
    return c_probs_out
  return dy_dt


def _checked_p0(p0, size_a, cl_k):
  p0 = numpy.asarray(p0, dtype=numpy.float64).ravel()
  if not (p0.size == size_a**cl_k and (0 <= p0).all() and (p0 <= 1).all()
          and abs(p0.sum() - 1) < 1e-10):
    raise ValueError('Parameter p0 is not a subsequence probability distribution.')
  return p0


def ode_integrate(*, tag, size_a, cl_k, p0, ts, odeint_kwargs=types.MappingProxyType({}),
                  debug=False):
  """scipy.integrate.odeint over the GPU right-hand side (markov_tapes.py:292-318)."""
  p0 = _checked_p0(p0, size_a, cl_k)
  dy_dt = get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k, debug=debug)
  return scipy.integrate.odeint(dy_dt, p0, ts, **odeint_kwargs)


def ode_integrate_ivp(*, tag, size_a, cl_k, p0, ts, ivp_kwargs=types.MappingProxyType({}),
                      debug=False):
  """scipy.integrate.solve_ivp over the GPU right-hand side, result shaped like odeint's
  (markov_tapes.py:321-354)."""
  p0 = _checked_p0(p0, size_a, cl_k)
  dy_dt = get_dy_dt(tag=tag, size_a=size_a, cl_k=cl_k, debug=debug)
  return scipy.integrate.solve_ivp(
      lambda t, y: dy_dt(y, t), (ts[0], ts[-1]), p0, t_eval=ts, **ivp_kwargs).y.T


### Additions

def register_rule_set(tag, size_a, rules):
  """Registers a rewrite-rule set under `tag` (see configs.random_rule_set)."""
  _lib.register_rules(tag, size_a, rules)


def register_program(tag, size_a, body):
  """Registers the problem stated by the Python function `body(tape)` under `tag`; `tape` offers
  get(data_tape, index), set(data_tape, index, symbol), choose(weights) and choose_value(pairs),
  the primitives of framework/gambit_macros.scm:99-125 (see programs.py).  Replaces editing
  framework/problems.scm and rebuilding the shared object."""
  from . import programs
  _lib.register_program(tag, size_a, programs.trace(body, size_a))


def model_stats(*, tag, cl_k):
  """Builds (or fetches) the device structure for (tag, cl_k) and returns its size facts."""
  model = u_lib.tapes_model(tag.encode(), cl_k)
  _lib.check(bool(model), 'tapes_model')
  stats = _lib.model_info(model)
  stats.update(_lib.model_timing(model))
  return stats


def _dop853_tableau():
  """SciPy's DOP853 coefficients, flattened in the order tapes_dop853_create expects."""
  from scipy.integrate._ivp import dop853_coefficients as dc
  parts = [dc.A, dc.B, dc.C, dc.E3, dc.E5, dc.D]
  flat = numpy.concatenate([numpy.asarray(x, dtype=numpy.float64).ravel() for x in parts])
  assert flat.size == 374
  return numpy.ascontiguousarray(flat)


def sequence_observable(size_a, cl_k, seq):
  """(offset, stride, count) of the strided sum that equals seq_prob(spd, seq) for len(seq) <= k."""
  if len(seq) > cl_k:
    raise ValueError('device observables need len(seq) <= cl_k; use seq_prob on fetched states')
  offset = 0
  for s in seq:
    offset = offset * size_a + int(s)
  return offset, size_a ** len(seq), size_a ** (cl_k - len(seq))


def ode_integrate_device(*, tag, size_a, cl_k, p0, ts, rtol=1e-3, atol=1e-6, max_step=numpy.inf,
                         first_step=None, observables=None, return_states=True, want_stats=False,
                         peer_group=None, entropy=False, eps=None):
  """DOP853 integration with the table resident in HBM (no per-stage host round trips).

  Follows scipy.integrate.solve_ivp(method='DOP853', t_eval=ts) step for step, so the result
  matches `ode_integrate_ivp(..., ivp_kwargs=dict(method='DOP853', rtol=..., atol=...))`.

  Returns the states at `ts` shaped like odeint's output ([len(ts), size_a**cl_k]) and/or, when
  `observables` (a list of symbol sequences of any length) is given, their probabilities at `ts`
  ([len(ts), len(observables)]) as `seq_prob` defines them (framework/markov_tapes.py:190-233;
  sequences longer than cl_k are extended with the Markov process parameters clipped at `eps`,
  default 1e-100), computed on the device; with entropy=True also `markov_entropy`
  (framework/markov_tapes.py:178-187) at `ts` ([len(ts)]).  With return_states=False only these
  numbers cross the host boundary.

  Several GPUs: `peer_group` is the parallel.PeerExchangeRhs of this rank's share of the problem,
  a device.DeviceModel(tag, cl_k, part=(rank, world)) - or `tag` names a share of a rule set made
  with parallel.split_rule_set; every rank calls this function with the same arguments otherwise
  and gets the same result (right-hand sides are evaluated by all ranks together, everything else
  is replicated).
  """
  p0 = _checked_p0(p0, size_a, cl_k)
  ts = numpy.asarray(ts, dtype=numpy.float64)
  if ts.ndim != 1 or ts.size < 2 or not ((numpy.diff(ts) > 0).all() or (numpy.diff(ts) < 0).all()):
    raise ValueError('ts must be a strictly monotonic sequence of at least two times')
  part = getattr(peer_group.model, 'part', None) if peer_group is not None else None
  if part is not None:
    if (peer_group.model.tag, peer_group.model.cl_k) != (tag, cl_k):
      raise ValueError('peer_group belongs to a different model')
    model = peer_group.model.handle
  else:
    model = u_lib.tapes_model(tag.encode(), cl_k)
    _lib.check(bool(model), 'tapes_model')
    if peer_group is not None and peer_group.model.handle != model:
      raise ValueError('peer_group belongs to a different model')
  tab = _dop853_tableau()
  solver = u_lib.tapes_dop853_create_peer(model, peer_group.group if peer_group is not None else None,
                                          tab.ctypes.data, numpy.ascontiguousarray(p0).ctypes.data,
                                          float(ts[0]), float(ts[-1]), float(rtol), float(atol),
                                          float(max_step) if numpy.isfinite(max_step) else -1.0,
                                          float(first_step) if first_step else -1.0)
  _lib.check(bool(solver), 'tapes_dop853_create')
  n = size_a ** cl_k
  obs = None
  if observables is not None:
    obs = _lib.pack_sequences(observables)
  states = numpy.empty((ts.size, n), dtype=numpy.float64) if return_states else None
  series = numpy.empty((ts.size, len(observables)), dtype=numpy.float64) if obs is not None else None
  entropies = numpy.empty(ts.size, dtype=numpy.float64) if entropy else None
  one = numpy.zeros(1, dtype=numpy.float64)
  forward = ts[-1] > ts[0]
  info = numpy.zeros(6, dtype=numpy.float64)
  try:
    done = 0  # number of requested times already written
    status = None
    while status is None:
      code = u_lib.tapes_dop853_step(solver)
      if code == 1:
        status = 0
      elif code < 0:
        _lib.check(code != -2, 'tapes_dop853_step')
        raise RuntimeError('Required step size is less than spacing between numbers.')
      u_lib.tapes_dop853_info(solver, info.ctypes.data)
      t = info[0]
      # every requested time in (t_old, t], the one equal to t included (solve_ivp's loop)
      if forward:
        upto = int(numpy.searchsorted(ts, t, side='right'))
      else:
        upto = int(ts.size - numpy.searchsorted(ts[::-1], t, side='left'))
      for i in range(done, upto):
        rc = u_lib.tapes_dop853_dense(solver, float(ts[i]))
        _lib.check(rc == 0, 'tapes_dop853_dense')
        if states is not None:
          _lib.check(u_lib.tapes_dop853_fetch(solver, 1, states[i].ctypes.data) == 0, 'tapes_dop853_fetch')
        if series is not None and len(observables):
          rc = u_lib.tapes_dop853_observe_sequences(solver, 1, len(observables), obs[0].ctypes.data, obs[1].ctypes.data,
                                                    1e-100 if eps is None else float(eps), series[i].ctypes.data)
          _lib.check(rc == 0, 'tapes_dop853_observe_sequences')
        if entropies is not None:
          _lib.check(u_lib.tapes_dop853_entropy(solver, 1, one.ctypes.data) == 0, 'tapes_dop853_entropy')
          entropies[i] = one[0]
      done = max(done, upto)
    u_lib.tapes_dop853_info(solver, info.ctypes.data)
  finally:
    u_lib.tapes_dop853_destroy(solver)
  if peer_group is not None:
    peer_group.check()
  out = tuple(x for x in (states, series, entropies) if x is not None)
  out = out[0] if len(out) == 1 else out
  if want_stats:
    return out, dict(nfev=int(info[3]), accepted=int(info[4]), rejected=int(info[5]), t=float(info[0]))
  return out


def check_table(spd, *, size_a=None, cl_k=None, eps_mpp=None, max_iterations=0, tolerance=1e-14):
  """Validates a subsequence-probability table on the GPU without a dense eigen-decomposition.

  `get_ctm_eigenvalue1_eigenspace` (framework/markov_tapes.py:133-175) builds the
  A^(k-1) x A^(k-1) context transfer matrix and calls numpy.linalg.eig on it, which stops being
  possible around 10^4 contexts.  The matrix has A non-zeros per row, so the same questions are
  answered by streaming kernels (csrc/validate.cu) for tables of any size that fits the device.

  Args:
    spd: array of shape [A]*k, or flat with `size_a` and `cl_k` given; numpy or a CUDA torch tensor.
    eps_mpp: the clip of `mpp_from_spd` (default 1e-100, as there).
    max_iterations: power-iteration steps v <- T v from the uniform vector (0 = skip).
    tolerance: stop the iteration when two successive vectors differ by less (2-norm).

  Returns a dict: `total` (sum of the table), `marginal_distance` (2-norm of last-axis minus
  first-axis marginal: the quantity the reference compares with its `eps`),
  `stationarity_residual` (|T pi - pi| for the context marginal pi), and when iterating
  `power_distance` (|v - pi|: large when pi is not the only stationary vector the uniform start
  converges to), `last_change`, `iterations`.
  """
  if eps_mpp is None:
    eps_mpp = 1e-100
  on_device = hasattr(spd, 'is_cuda') and spd.is_cuda
  if size_a is None or cl_k is None:
    size_a, cl_k = int(spd.shape[0]), int(spd.ndim if not hasattr(spd, 'dim') else spd.dim())
  n = size_a ** cl_k
  if on_device:
    import torch
    if spd.dtype != torch.float64 or not spd.is_contiguous() or spd.numel() != n:
      raise ValueError(f'expected {n} contiguous float64 entries')
    torch.cuda.current_stream().synchronize()
    ptr, keep = spd.data_ptr(), spd
  else:
    keep = numpy.ascontiguousarray(numpy.asarray(spd, dtype=numpy.float64).ravel())
    if keep.size != n:
      raise ValueError(f'expected {n} entries, got {keep.size}')
    ptr = keep.ctypes.data
  out = numpy.zeros(6, dtype=numpy.float64)
  rc = u_lib.tapes_check_table(size_a, cl_k, ptr, 1 if on_device else 0, float(eps_mpp), int(max_iterations),
                               float(tolerance), out.ctypes.data)
  _lib.check(rc == 0, 'tapes_check_table')
  del keep
  result = dict(total=float(out[0]), marginal_distance=float(out[1]), stationarity_residual=float(out[2]))
  if max_iterations > 0:
    result.update(power_distance=float(out[3]), last_change=float(out[4]), iterations=int(out[5]))
  return result


def monte_carlo(*, tag, size_a, cl_k, ts, p0=None, tape0=None, n_sites=1 << 20, events_per_substep=None, seed=0,
                return_tape=False):
  """Monte-Carlo estimate of the subsequence table over time: the problem's program is run at
  random positions of one long ring tape, every site being visited at rate 1, and the length-cl_k
  window frequencies are read off at the times `ts` (csrc/montecarlo.cu).  An independent check of
  the closure behind `ode_integrate*`: where the closure is exact (e.g. ex1) the two agree within
  the statistical error ~ n_sites**-0.5, elsewhere the difference is what the closure neglects.
  The reference has such a simulation for the ferromagnet only
  (examples/ex2_ferromagnet_mc.py:46-122).

  The ring starts as `tape0` (symbols, one per site) or is sampled from the table `p0`.  Events
  happen `events_per_substep` at a time (default n_sites / 1000, i.e. dt = 0.001).  Returns
  [len(ts), size_a**cl_k] frequencies shaped like `ode_integrate`'s output (and the final ring with
  return_tape=True)."""
  ts = numpy.asarray(ts, dtype=numpy.float64)
  if ts.ndim != 1 or ts.size < 1 or (numpy.diff(ts) < 0).any() or ts[0] < 0:
    raise ValueError('ts must be a non-decreasing sequence of non-negative times')
  if tape0 is None:
    if p0 is None:
      raise ValueError('give the initial ring (tape0) or a table to sample it from (p0)')
    tape0 = _lib.sample_ring(size_a, cl_k, _checked_p0(p0, size_a, cl_k), int(n_sites), int(seed))
  tape0 = numpy.ascontiguousarray(numpy.asarray(tape0, dtype=numpy.uint8).ravel())
  n_sites = tape0.size
  if events_per_substep is None:
    events_per_substep = max(1, min(n_sites // 1000, (1 << 20) - 1))  # the library takes fewer than 2^20
  mc = u_lib.tapes_mc_create(tag.encode(), n_sites, tape0.ctypes.data, int(events_per_substep), int(seed))
  _lib.check(bool(mc), 'tapes_mc_create')
  out = numpy.zeros((ts.size, size_a ** cl_k), dtype=numpy.float64)
  counts = numpy.zeros(size_a ** cl_k, dtype=numpy.int64)
  done = 0
  try:
    for i, t in enumerate(ts):
      target = int(round(t * n_sites / events_per_substep))
      _lib.check(u_lib.tapes_mc_run(mc, target - done) == 0, 'tapes_mc_run')
      done = target
      _lib.check(u_lib.tapes_mc_window_counts(mc, cl_k, counts.ctypes.data) == 0, 'tapes_mc_window_counts')
      out[i] = counts / float(n_sites)
    if return_tape:
      tape = numpy.zeros(n_sites, dtype=numpy.uint8)
      _lib.check(u_lib.tapes_mc_fetch(mc, tape.ctypes.data) == 0, 'tapes_mc_fetch')
      return out, tape
  finally:
    u_lib.tapes_mc_destroy(mc)
  return out


def ferromagnet_mc_inputs(*, n_trials, chain_length, n_steps, sites_per_pair, trials_per_step, beta=1.0, J=1.0, h=-0.25,
                          seed_offset=1000, first_trial=0):
  """Everything random or parameter-dependent of the reference's ferromagnet Monte Carlo
  (examples/ex2_ferromagnet_mc.py), drawn exactly as the script draws it: per trial
  numpy.random.RandomState(seed = trial + seed_offset) (172), the initial chain from one uniform
  vector (175-176), then per time step `randint` for the sites and `uniform` for the acceptance
  tests (93-94).  Returns (chain0 [T, N] uint8, sites [T, S-1, M] int32, uniforms [T, S-1, M] float64,
  accept [3, 2]): accept[c, s] is the flip probability of a spin s with c equal neighbours (101-113)."""
  chain0 = numpy.zeros((n_trials, chain_length), dtype=numpy.uint8)
  sites = numpy.zeros((n_trials, max(n_steps - 1, 0), trials_per_step), dtype=numpy.int32)
  uniforms = numpy.zeros((n_trials, max(n_steps - 1, 0), trials_per_step), dtype=numpy.float64)
  for t in range(n_trials):
    rng = numpy.random.RandomState(seed=first_trial + t + seed_offset)
    pair_positions = rng.uniform(0, 1, size=chain_length) < 1 / sites_per_pair
    chain0[t] = (pair_positions | numpy.roll(pair_positions, 1)).astype(numpy.uint8)
    for nt in range(n_steps - 1):
      sites[t, nt] = rng.randint(0, chain_length, size=trials_per_step)
      uniforms[t, nt] = rng.uniform(0, 1, size=trials_per_step)
  beta_j, beta_h = beta * J, beta * h
  accept = numpy.zeros((3, 2), dtype=numpy.float64)
  for equal in range(3):
    energy_change = 2 * (equal - (2 - equal))  # neighbour energy after minus before, in units of J
    for spin in (0, 1):
      e_j = numpy.exp(-beta_j * (energy_change + 4))
      e_h = numpy.exp(-2 * beta_h * spin) if h > 0 else numpy.exp(+2 * beta_h * (1 - spin))
      accept[equal, spin] = e_j * e_h
  return chain0, sites, uniforms, accept


def ferromagnet_monte_carlo(*, n_trials=100, chain_length=50000, n_steps=4000, sites_per_pair=250, trials_per_step=None,
                            beta=1.0, J=1.0, h=-0.25, seed_offset=1000, batch=10):
  """The reference's Monte-Carlo experiment on the ferromagnetic chain
  (examples/ex2_ferromagnet_mc.py:33-45, 167-191) on the GPU, with the script's defaults: returns
  `chain_counts` [n_trials, n_steps, 6], the number of up-spin islands of length 1..5 at every time
  step of every trial - the array the script stores as `chain_counts` in
  ferromagnet_mc_chain_counts.npz.  The random numbers are the script's own (ferromagnet_mc_inputs),
  so the counts are the script's counts, number for number.  One thread block per trial, chain in
  shared memory (csrc/montecarlo.cu); trials go to the device `batch` at a time."""
  if trials_per_step is None:
    trials_per_step = chain_length // 100
  counts = numpy.zeros((n_trials, n_steps, 6), dtype=numpy.float64)
  for first in range(0, n_trials, batch):
    n = min(batch, n_trials - first)
    chain0, sites, uniforms, accept = ferromagnet_mc_inputs(
        n_trials=n, chain_length=chain_length, n_steps=n_steps, sites_per_pair=sites_per_pair,
        trials_per_step=trials_per_step, beta=beta, J=J, h=h, seed_offset=seed_offset, first_trial=first)
    part = numpy.zeros((n, n_steps, 6), dtype=numpy.float64)
    rc = u_lib.tapes_mc_ferromagnet_chains(n, chain_length, n_steps, trials_per_step, chain0.ctypes.data, sites.ctypes.data,
                                           uniforms.ctypes.data, accept.ctypes.data, part.ctypes.data)
    _lib.check(rc == 0, 'tapes_mc_ferromagnet_chains')
    counts[first:first + n] = part
  return counts


def _run_validation():
  fn_dy_dt = get_dy_dt(tag='__canary_problem_radioactive_decay', size_a=2, cl_k=3, debug=False)
  observed = fn_dy_dt(numpy.full([8], fill_value=0.125, dtype=numpy.float64), 0.0).tolist()
  expected = [0.375, 0.125, 0.125, -0.125, 0.125, -0.125, -0.125, -0.375]
  if expected != observed:
    raise RuntimeError('Load-time validation problem failed to produce the expected result.')


init_gambit()
_run_validation()
