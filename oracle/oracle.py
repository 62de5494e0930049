"""ctypes front end of the CPU oracle (oracle/tape_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

The call surface mirrors framework/markov_tapes.py:259-289 (`get_dy_dt`) of the reference so
that the same SciPy steppers can drive it (framework/markov_tapes.py:292-354).
"""

import ctypes
import os
import subprocess

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'liboracle.so')

LITERAL = 0
MERGED = 1


def build(force=False):
  """Compiles the oracle with the committed Makefile (g++)."""
  src = os.path.join(_HERE, 'tape_oracle.cpp')
  if (force or not os.path.exists(_LIB_PATH)
      or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
    subprocess.check_call(['make', '-C', _HERE, '-s'])
  return _LIB_PATH


_lib = None


def lib():
  global _lib
  if _lib is None:
    if not os.path.exists(_LIB_PATH):
      build()
    _lib = ctypes.CDLL(_LIB_PATH)
    _lib.oracle_compute_dy_dt.restype = ctypes.c_int
    _lib.oracle_compute_dy_dt.argtypes = [
        ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p]
    _lib.oracle_terms.restype = ctypes.c_int64
    _lib.oracle_terms.argtypes = [
        ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    _lib.oracle_worlds.restype = ctypes.c_int64
    _lib.oracle_worlds.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p,
                                   ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    _lib.oracle_alphabet_size.restype = ctypes.c_int64
    _lib.oracle_alphabet_size.argtypes = [ctypes.c_char_p]
    _lib.oracle_register_rules.restype = ctypes.c_int
    _lib.oracle_register_rules.argtypes = [
        ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 7
    _lib.oracle_register_program.restype = ctypes.c_int
    _lib.oracle_register_program.argtypes = (
        [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 6
        + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p])
    _lib.oracle_last_error.restype = ctypes.c_char_p
  return _lib


def last_error():
  return lib().oracle_last_error().decode()


def alphabet_size(tag):
  return int(lib().oracle_alphabet_size(tag.encode()))


def compute_dy_dt(tag, cl_k, probs, mode=LITERAL, want_counters=False):
  """One right-hand side.  Returns dy/dt (and the counters dict if requested)."""
  p = numpy.ascontiguousarray(numpy.asarray(probs, dtype=numpy.float64).ravel())
  size_a = alphabet_size(tag)
  if size_a < 0:
    raise KeyError(f'unknown problem tag {tag!r}')
  if p.size != size_a ** cl_k:
    raise ValueError(f'probability-array should have size {size_a ** cl_k}, observed: {p.size}')
  out = numpy.zeros_like(p)
  counters = numpy.zeros(4, dtype=numpy.int64)
  rc = lib().oracle_compute_dy_dt(tag.encode(), cl_k, mode, p.ctypes.data, out.ctypes.data,
                                  counters.ctypes.data)
  if rc != 0:
    raise RuntimeError(last_error())
  if want_counters:
    return out, dict(worlds=int(counters[0]), ext_nodes=int(counters[1]),
                     acc_calls=int(counters[2]), leaf_worlds=int(counters[3]))
  return out


def terms(tag, cl_k, probs, mode=LITERAL):
  """All accumulated flux terms (src, dst, w) in the oracle's accumulation order."""
  p = numpy.ascontiguousarray(numpy.asarray(probs, dtype=numpy.float64).ravel())
  n = lib().oracle_terms(tag.encode(), cl_k, mode, p.ctypes.data, 0, None, None, None)
  if n < 0:
    raise RuntimeError(last_error())
  src = numpy.zeros(n, dtype=numpy.int64)
  dst = numpy.zeros(n, dtype=numpy.int64)
  w = numpy.zeros(n, dtype=numpy.float64)
  lib().oracle_terms(tag.encode(), cl_k, mode, p.ctypes.data, n, src.ctypes.data,
                     dst.ctypes.data, w.ctypes.data)
  return src, dst, w


def worlds(tag, cl_k, probs):
  """Leaf worlds at `probs`: (prob[n], info[n, 6]) with info = plen, p_orig, p_adj, dlen, d_orig,
  d_adj (see oracle_worlds in tape_oracle.cpp)."""
  p = numpy.ascontiguousarray(numpy.asarray(probs, dtype=numpy.float64).ravel())
  n = lib().oracle_worlds(tag.encode(), cl_k, p.ctypes.data, 0, None, None)
  if n < 0:
    raise RuntimeError(last_error())
  prob = numpy.zeros(n, dtype=numpy.float64)
  info = numpy.zeros((n, 6), dtype=numpy.int64)
  lib().oracle_worlds(tag.encode(), cl_k, p.ctypes.data, n, prob.ctypes.data, info.ctypes.data)
  return prob, info


def register_rules(tag, size_a, rules):
  """Registers a rule-table problem; `rules` is the dict produced by rule-table generators
  (keys tape, span, catalyst, pattern[R,4], repl[R,4], rate, select_weight)."""
  arrs = _rule_arrays(rules)
  rc = lib().oracle_register_rules(tag.encode(), size_a, arrs[0].shape[0],
                                   *[a.ctypes.data for a in arrs])
  if rc != 0:
    raise RuntimeError(last_error())


def register_program(tag, size_a, tree):
  """Registers a problem given as a decision tree (dict of arrays: kind, a, b, c, first_child,
  first_weight per node, child, weight; see oracle_register_program in tape_oracle.cpp)."""
  i32 = lambda x: numpy.ascontiguousarray(numpy.asarray(x, dtype=numpy.int32))
  cols = [i32(tree[key]) for key in ('kind', 'a', 'b', 'c', 'first_child', 'first_weight')]
  child = i32(tree['child'])
  weight = numpy.ascontiguousarray(numpy.asarray(tree['weight'], dtype=numpy.float64))
  rc = lib().oracle_register_program(tag.encode(), size_a, cols[0].size, *[col.ctypes.data for col in cols],
                                     child.size, child.ctypes.data, weight.size, weight.ctypes.data)
  if rc != 0:
    raise RuntimeError(last_error())


def _rule_arrays(rules):
  i32 = lambda x: numpy.ascontiguousarray(numpy.asarray(x, dtype=numpy.int32))
  f64 = lambda x: numpy.ascontiguousarray(numpy.asarray(x, dtype=numpy.float64))
  pattern = i32(rules['pattern']).reshape(-1, 4)
  repl = i32(rules['repl']).reshape(-1, 4)
  return (i32(rules['tape']), i32(rules['span']), i32(rules['catalyst']), pattern, repl,
          f64(rules['rate']), f64(rules['select_weight']))


def get_dy_dt(*, tag, size_a, cl_k, mode=LITERAL):
  """Same closure shape as the reference's get_dy_dt (framework/markov_tapes.py:259-289)."""
  expected_size = size_a ** cl_k
  if alphabet_size(tag) != size_a:
    raise ValueError(f'alphabet size mismatch for {tag!r}')

  def dy_dt(a_probs_in, t):
    del t
    c_probs_in = numpy.asarray(a_probs_in, dtype=numpy.float64).ravel()
    if c_probs_in.size != expected_size:
      raise ValueError(f'probability-array should have size {expected_size}, '
                       f'observed: {c_probs_in.size}')
    return compute_dy_dt(tag, cl_k, c_probs_in, mode=mode)
  return dy_dt
