"""ctypes binding of tapes_py_interface.so (C ABI: include/tapes_b200.h).

Loading fails loudly when the library has not been built; there is no CPU fallback.
"""

import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'tapes_py_interface.so')

# every symbol include/tapes_b200.h declares
SYMBOLS = (
    'setup_gambit', 'cleanup_gambit', 'c_register_problems', 'c_compute_dy_dt',
    'tapes_last_error', 'tapes_clear_error', 'tapes_alphabet_size', 'tapes_register_rules',
    'tapes_model', 'tapes_release_model', 'tapes_rhs_device', 'tapes_weights_device', 'tapes_flux_rows_device', 'tapes_rhs_profile', 'tapes_sync', 'tapes_model_info',
    'tapes_model_set', 'tapes_model_timing', 'tapes_export_csr', 'tapes_export_node_weights', 'tapes_rule_table',
    'tapes_peer_alloc', 'tapes_peer_open', 'tapes_peer_close', 'tapes_peer_free', 'tapes_peer_group_create',
    'tapes_peer_group_destroy', 'tapes_peer_rhs', 'tapes_peer_group_error', 'tapes_dop853_create_peer',
    'tapes_check_table', 'tapes_model_part', 'tapes_rule_parts', 'tapes_register_program',
    'tapes_mc_create', 'tapes_mc_destroy', 'tapes_mc_run', 'tapes_mc_window_counts', 'tapes_mc_fetch',
    'tapes_mc_sample_ring', 'tapes_program_tree', 'tapes_observe_sequences', 'tapes_dop853_observe_sequences',
    'tapes_markov_entropy', 'tapes_dop853_entropy', 'tapes_host_alloc', 'tapes_host_free', 'tapes_mc_ferromagnet_chains',
)

_lib = None


def load():
  """Loads the shared library and declares its signatures."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise OSError(
        f'{LIB_PATH} is missing: build it with '
        '`python -m chemical_kinetics_and_program_execution_b200.build_ext` '
        '(or __graft_entry__.build()); this package has no CPU fallback.')
  lib = ctypes.CDLL(LIB_PATH)
  vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
  lib.setup_gambit.restype = vp
  lib.setup_gambit.argtypes = []
  lib.cleanup_gambit.restype = None
  lib.cleanup_gambit.argtypes = [vp]
  lib.c_register_problems.restype = i64
  lib.c_register_problems.argtypes = [i64]
  lib.c_compute_dy_dt.restype = None
  lib.c_compute_dy_dt.argtypes = [vp, i64, i64, vp, vp]
  lib.tapes_last_error.restype = ctypes.c_char_p
  lib.tapes_last_error.argtypes = []
  lib.tapes_clear_error.restype = None
  lib.tapes_clear_error.argtypes = []
  lib.tapes_alphabet_size.restype = i64
  lib.tapes_alphabet_size.argtypes = [ctypes.c_char_p]
  lib.tapes_register_rules.restype = i32
  lib.tapes_register_rules.argtypes = [ctypes.c_char_p, i64, i64] + [vp] * 7
  lib.tapes_register_program.restype = i32
  lib.tapes_register_program.argtypes = [ctypes.c_char_p, i64, i64] + [vp] * 6 + [i64, vp, i64, vp]
  lib.tapes_mc_create.restype = vp
  lib.tapes_mc_create.argtypes = [ctypes.c_char_p, i64, vp, i64, ctypes.c_uint64]
  lib.tapes_mc_destroy.restype = None
  lib.tapes_mc_destroy.argtypes = [vp]
  lib.tapes_mc_run.restype = i32
  lib.tapes_mc_run.argtypes = [vp, i64]
  lib.tapes_mc_window_counts.restype = i32
  lib.tapes_mc_window_counts.argtypes = [vp, i64, vp]
  lib.tapes_mc_fetch.restype = i32
  lib.tapes_mc_fetch.argtypes = [vp, vp]
  lib.tapes_mc_sample_ring.restype = i32
  lib.tapes_mc_sample_ring.argtypes = [i64, i64, vp, i64, ctypes.c_uint64, vp]
  lib.tapes_program_tree.restype = i64
  lib.tapes_program_tree.argtypes = [ctypes.c_char_p] + [vp] * 9
  lib.tapes_model.restype = vp
  lib.tapes_model.argtypes = [ctypes.c_char_p, i64]
  lib.tapes_model_part.restype = vp
  lib.tapes_model_part.argtypes = [ctypes.c_char_p, i64, i64, i64]
  lib.tapes_rule_parts.restype = i64
  lib.tapes_rule_parts.argtypes = [ctypes.c_char_p, i64, i64, vp, vp]
  lib.tapes_release_model.restype = i32
  lib.tapes_release_model.argtypes = [ctypes.c_char_p, i64]
  lib.tapes_rhs_device.restype = i32
  lib.tapes_rhs_device.argtypes = [vp, vp, vp, vp]
  lib.tapes_weights_device.restype = i32
  lib.tapes_weights_device.argtypes = [vp, vp, vp]
  lib.tapes_flux_rows_device.restype = i32
  lib.tapes_flux_rows_device.argtypes = [vp, vp, i64, i64, vp]
  lib.tapes_peer_alloc.restype = vp
  lib.tapes_peer_alloc.argtypes = [i64, vp]
  lib.tapes_peer_open.restype = vp
  lib.tapes_peer_open.argtypes = [vp]
  lib.tapes_peer_close.restype = i32
  lib.tapes_peer_close.argtypes = [vp]
  lib.tapes_peer_free.restype = i32
  lib.tapes_peer_free.argtypes = [vp]
  lib.tapes_peer_group_create.restype = vp
  lib.tapes_peer_group_create.argtypes = [i32, i32, i64, i32, vp, vp, vp]
  lib.tapes_peer_group_destroy.restype = None
  lib.tapes_peer_group_destroy.argtypes = [vp]
  lib.tapes_peer_rhs.restype = i32
  lib.tapes_peer_rhs.argtypes = [vp, vp, vp, vp]
  lib.tapes_peer_group_error.restype = i32
  lib.tapes_peer_group_error.argtypes = [vp]
  lib.tapes_dop853_create_peer.restype = vp
  lib.tapes_dop853_create_peer.argtypes = [vp, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, dbl]
  lib.tapes_check_table.restype = i32
  lib.tapes_check_table.argtypes = [i64, i64, vp, i32, dbl, i64, dbl, vp]
  lib.tapes_rhs_profile.restype = i32
  lib.tapes_rhs_profile.argtypes = [vp, vp, vp, vp, vp, i32]
  lib.tapes_sync.restype = i32
  lib.tapes_sync.argtypes = [vp]
  lib.tapes_model_info.restype = i32
  lib.tapes_model_info.argtypes = [vp, vp, i32]
  lib.tapes_model_set.restype = i32
  lib.tapes_model_set.argtypes = [vp, ctypes.c_char_p, i64]
  lib.tapes_model_timing.restype = i32
  lib.tapes_model_timing.argtypes = [vp, vp, i32]
  lib.tapes_export_csr.restype = i32
  lib.tapes_export_csr.argtypes = [vp, vp, vp]
  lib.tapes_export_node_weights.restype = i32
  lib.tapes_export_node_weights.argtypes = [vp, vp]
  lib.tapes_rule_table.restype = i64
  lib.tapes_rule_table.argtypes = [ctypes.c_char_p, i64] + [vp] * 11
  lib.tapes_dop853_create.restype = vp
  lib.tapes_dop853_create.argtypes = [vp, vp, vp, dbl, dbl, dbl, dbl, dbl, dbl]
  lib.tapes_dop853_destroy.restype = None
  lib.tapes_dop853_destroy.argtypes = [vp]
  lib.tapes_dop853_step.restype = i32
  lib.tapes_dop853_step.argtypes = [vp]
  lib.tapes_dop853_dense.restype = i32
  lib.tapes_dop853_dense.argtypes = [vp, dbl]
  lib.tapes_dop853_fetch.restype = i32
  lib.tapes_dop853_fetch.argtypes = [vp, i32, vp]
  lib.tapes_dop853_observe.restype = i32
  lib.tapes_dop853_observe.argtypes = [vp, i32, vp, vp, vp, i64, vp]
  lib.tapes_dop853_info.restype = i32
  lib.tapes_dop853_info.argtypes = [vp, vp]
  lib.tapes_observe.restype = i32
  lib.tapes_observe.argtypes = [vp, vp, vp, vp, vp, i64, vp]
  lib.tapes_observe_sequences.restype = i32
  lib.tapes_observe_sequences.argtypes = [vp, vp, i64, vp, vp, dbl, vp]
  lib.tapes_dop853_observe_sequences.restype = i32
  lib.tapes_dop853_observe_sequences.argtypes = [vp, i32, i64, vp, vp, dbl, vp]
  lib.tapes_markov_entropy.restype = i32
  lib.tapes_markov_entropy.argtypes = [vp, vp, vp]
  lib.tapes_dop853_entropy.restype = i32
  lib.tapes_dop853_entropy.argtypes = [vp, i32, vp]
  lib.tapes_mc_ferromagnet_chains.restype = i32
  lib.tapes_mc_ferromagnet_chains.argtypes = [i64, i64, i64, i64, vp, vp, vp, vp, vp]
  lib.tapes_host_alloc.restype = vp
  lib.tapes_host_alloc.argtypes = [i64]
  lib.tapes_host_free.restype = i32
  lib.tapes_host_free.argtypes = [vp]
  _lib = lib
  return lib


def last_error():
  return load().tapes_last_error().decode()


def check(ok, what):
  """Raises RuntimeError carrying the library's error message when `ok` is false."""
  if not ok:
    msg = last_error()
    load().tapes_clear_error()
    raise RuntimeError(f'{what}: {msg}' if msg else what)


MODEL_INFO_FIELDS = (
    'n_states', 'n_nodes', 'nnz', 'n_flux_rules', 'n_levels', 'launches_per_rhs', 'n_terms',
    'n_sum_nodes', 'worlds_walked', 'leaf_worlds', 'seeds', 'hash_inserts', 'hash_unique',
    'alphabet', 'cl_k', 'spmv_lanes_per_row', 'flux_format', 'n_slices', 'slice_words', 'runs',
    'run_entries', 'column_entries', 'column_slots', 'min_run_lanes', 'level_unroll',
    'irregular_levels', 'left_parents', 'flux_unroll', 'owned_parents', 'deferred_groups', 'structures',
    'interleaved_levels', 'plane_groups', 'ratio_tables', 'nnz_stored', 'materialize_right',
    'hash_retries', 'distinct_table_reads')


def model_info(model):
  buf = numpy.zeros(len(MODEL_INFO_FIELDS), dtype=numpy.int64)
  n = load().tapes_model_info(model, buf.ctypes.data, buf.size)
  return {name: int(buf[i]) for i, name in enumerate(MODEL_INFO_FIELDS[:n])}


def model_timing(model):
  buf = numpy.zeros(5, dtype=numpy.float64)
  load().tapes_model_timing(model, buf.ctypes.data, 5)
  return dict(host_enumerate_ms=float(buf[0]), device_expand_ms=float(buf[1]),
              device_csr_ms=float(buf[2]), device_slices_ms=float(buf[3]), expand_alloc_ms=float(buf[4]))


class PinnedResults:
  """Result arrays of get_dy_dt for large tables: NumPy arrays over page-locked memory, so that the
  device-to-host copy is plain DMA and no fresh 8 * n bytes are zero-filled per call.  A buffer goes
  back to the pool when the array that wraps it is garbage-collected (the SciPy steppers copy the
  returned values into their own stage arrays and drop the array at once); at most `limit` buffers
  exist, and when all of them are still referenced the caller gets an ordinary NumPy array."""

  def __init__(self, n_doubles, limit=4):
    self.n, self.limit = int(n_doubles), int(limit)
    self.free, self.alive = [], 0

  def take(self):
    import weakref
    if self.free:
      ptr = self.free.pop()
    elif self.alive < self.limit:
      ptr = load().tapes_host_alloc(8 * self.n)
      if not ptr:
        load().tapes_clear_error()
        return None
      self.alive += 1
    else:
      return None
    buf = (ctypes.c_double * self.n).from_address(ptr)
    arr = numpy.frombuffer(buf, dtype=numpy.float64, count=self.n)
    weakref.finalize(buf, self.free.append, ptr)  # `arr` (and every view of it) keeps `buf` alive
    return arr

  def release(self):
    """Frees the idle buffers (buffers still wrapped by live arrays stay valid)."""
    while self.free:
      load().tapes_host_free(self.free.pop())
      self.alive -= 1


def pack_sequences(seqs):
  """(seq_ptr[int64], symbols[int32]) of a list of symbol sequences, the form tapes_observe_sequences takes."""
  seq_ptr = numpy.zeros(len(seqs) + 1, dtype=numpy.int64)
  seq_ptr[1:] = numpy.cumsum([len(s) for s in seqs])
  symbols = numpy.ascontiguousarray(numpy.array([int(x) for s in seqs for x in s] or [0], dtype=numpy.int32))
  return seq_ptr, symbols


def rule_table(tag, cl_k):
  """Host-only flux-rule table of (tag, cl_k) as a dict of numpy arrays."""
  lib = load()
  n_steps = ctypes.c_int64(0)
  stats = numpy.zeros(2, dtype=numpy.int64)
  n_rules = lib.tapes_rule_table(tag.encode(), cl_k, ctypes.addressof(n_steps), None, None, None,
                                 None, None, None, None, None, None, stats.ctypes.data)
  check(n_rules >= 0, 'tapes_rule_table')
  ns = n_steps.value
  out = dict(
      rule_ptr=numpy.zeros(n_rules + 1, dtype=numpy.int64),
      step_kind=numpy.zeros(ns, dtype=numpy.int32), step_len=numpy.zeros(ns, dtype=numpy.int32),
      step_long=numpy.zeros(ns, dtype=numpy.int64), step_short=numpy.zeros(ns, dtype=numpy.int64),
      step_prob=numpy.zeros(ns, dtype=numpy.float64),
      seed_len=numpy.zeros((n_rules, 2), dtype=numpy.int32),
      seed_orig=numpy.zeros((n_rules, 2), dtype=numpy.uint64),
      seed_adj=numpy.zeros((n_rules, 2), dtype=numpy.uint64))
  n2 = lib.tapes_rule_table(
      tag.encode(), cl_k, ctypes.addressof(n_steps), out['rule_ptr'].ctypes.data,
      out['step_kind'].ctypes.data, out['step_len'].ctypes.data, out['step_long'].ctypes.data,
      out['step_short'].ctypes.data, out['step_prob'].ctypes.data, out['seed_len'].ctypes.data,
      out['seed_orig'].ctypes.data, out['seed_adj'].ctypes.data, stats.ctypes.data)
  check(n2 == n_rules, 'tapes_rule_table')
  out['worlds_walked'] = int(stats[0])
  out['leaf_worlds'] = int(stats[1])
  return out


def rule_parts(tag, cl_k, n_parts):
  """Host-only: (owner[int32], cost[float64]) of every flux rule of (tag, cl_k) when the problem
  is dealt to n_parts ranks (the dealing tapes_model_part uses)."""
  lib = load()
  n_rules = lib.tapes_rule_parts(tag.encode(), cl_k, n_parts, None, None)
  check(n_rules >= 0, 'tapes_rule_parts')
  owner = numpy.zeros(n_rules, dtype=numpy.int32)
  cost = numpy.zeros(n_rules, dtype=numpy.float64)
  check(lib.tapes_rule_parts(tag.encode(), cl_k, n_parts, owner.ctypes.data, cost.ctypes.data) == n_rules,
        'tapes_rule_parts')
  return owner, cost


def program_tree(tag):
  """Host-only: the decision tree of the body registered under `tag` (dict of arrays, the form
  register_program takes)."""
  lib = load()
  sizes = numpy.zeros(3, dtype=numpy.int64)
  n = lib.tapes_program_tree(tag.encode(), sizes.ctypes.data, *([None] * 8))
  check(n >= 0, 'tapes_program_tree')
  tree = {key: numpy.zeros(sizes[0], dtype=numpy.int32) for key in ('kind', 'a', 'b', 'c', 'first_child', 'first_weight')}
  tree['child'] = numpy.zeros(sizes[1], dtype=numpy.int32)
  tree['weight'] = numpy.zeros(sizes[2], dtype=numpy.float64)
  n2 = lib.tapes_program_tree(tag.encode(), sizes.ctypes.data, *[tree[key].ctypes.data for key in (
      'kind', 'a', 'b', 'c', 'first_child', 'first_weight', 'child', 'weight')])
  check(n2 == n, 'tapes_program_tree')
  return tree


def sample_ring(size_a, cl_k, table, n_sites, seed):
  """Host-only: a ring of n_sites symbols whose length-cl_k window statistics follow `table`."""
  table = numpy.ascontiguousarray(numpy.asarray(table, dtype=numpy.float64).ravel())
  if table.size != size_a ** cl_k:
    raise ValueError(f'expected {size_a ** cl_k} table entries, got {table.size}')
  tape = numpy.zeros(n_sites, dtype=numpy.uint8)
  rc = load().tapes_mc_sample_ring(size_a, cl_k, table.ctypes.data, n_sites, seed, tape.ctypes.data)
  check(rc == 0, 'tapes_mc_sample_ring')
  return tape


def register_program(tag, size_a, tree):
  """Registers a program tree (dict of arrays as produced by programs.trace) under `tag`."""
  i32 = lambda x: numpy.ascontiguousarray(numpy.asarray(x, dtype=numpy.int32))
  cols = [i32(tree[key]) for key in ('kind', 'a', 'b', 'c', 'first_child', 'first_weight')]
  child = i32(tree['child'])
  weight = numpy.ascontiguousarray(numpy.asarray(tree['weight'], dtype=numpy.float64))
  rc = load().tapes_register_program(tag.encode(), size_a, cols[0].size, *[col.ctypes.data for col in cols],
                                     child.size, child.ctypes.data, weight.size, weight.ctypes.data)
  check(rc == 0, 'tapes_register_program')


def register_rules(tag, size_a, rules):
  """Registers a rewrite-rule set (dict with keys tape, span, catalyst, pattern[R,4], repl[R,4],
  rate, select_weight) under `tag`."""
  i32 = lambda x: numpy.ascontiguousarray(numpy.asarray(x, dtype=numpy.int32))
  f64 = lambda x: numpy.ascontiguousarray(numpy.asarray(x, dtype=numpy.float64))
  arrs = (i32(rules['tape']), i32(rules['span']), i32(rules['catalyst']),
          i32(rules['pattern']).reshape(-1, 4), i32(rules['repl']).reshape(-1, 4),
          f64(rules['rate']), f64(rules['select_weight']))
  rc = load().tapes_register_rules(tag.encode(), size_a, arrs[0].shape[0],
                                   *[a.ctypes.data for a in arrs])
  check(rc == 0, 'tapes_register_rules')
