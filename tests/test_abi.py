"""The C-ABI library loads without a GPU and exports every symbol include/tapes_b200.h declares."""

import ctypes
import os
import re

from conftest import ROOT
from chemical_kinetics_and_program_execution_b200 import _lib


def declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'tapes_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return re.findall(r'\b(\w+)\s*\(', ' '.join(l for l in text.splitlines() if not l.startswith('#')))


def test_header_symbols_are_exported():
  names = [n for n in declared_symbols() if n not in ('defined',)]
  assert 'c_compute_dy_dt' in names and 'setup_gambit' in names
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for name in names:
    assert hasattr(lib, name), name
  assert set(_lib.SYMBOLS) <= set(names)


def test_reference_symbols_have_reference_signatures():
  # framework/markov_tapes.py:43-56
  lib = _lib.load()
  assert lib.c_register_problems.argtypes == [ctypes.c_int64]
  assert len(lib.c_compute_dy_dt.argtypes) == 5


def reference_binding():
  """The ctypes declarations of the unmodified reference module, verbatim
  (framework/markov_tapes.py:38-56), against this repo's library."""
  u_lib = ctypes.CDLL(_lib.LIB_PATH)

  u_lib.setup_gambit.restype = ctypes.c_void_p
  u_lib.setup_gambit.argtypes = []

  u_lib.cleanup_gambit.restype = None
  u_lib.cleanup_gambit.argtypes = [ctypes.c_void_p]

  u_lib.c_register_problems.restype = ctypes.c_int64
  u_lib.c_register_problems.argtypes = [ctypes.c_int64]

  u_lib.c_compute_dy_dt.restype = None
  u_lib.c_compute_dy_dt.argtypes = [ctypes.c_void_p,
                                    ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p]
  return u_lib


def reference_dy_dt(u_lib, tag, cl_k, probs):
  """The call sequence of the reference's dy_dt closure (framework/markov_tapes.py:272-288): a
  zero-filled result buffer goes in, nothing is checked afterwards."""
  import numpy
  a_tag = numpy.frombuffer(tag.encode() + b'\x00', dtype=numpy.uint8)
  c_probs_in = numpy.asarray(probs, dtype=numpy.float64).ravel()
  c_probs_out = numpy.zeros_like(c_probs_in)
  u_lib.c_compute_dy_dt(a_tag.__array_interface__['data'][0], cl_k, 0,
                        c_probs_in.__array_interface__['data'][0],
                        c_probs_out.__array_interface__['data'][0])
  return c_probs_out


def test_failed_call_cannot_pass_for_a_steady_state(capfd):
  """The reference never returns from a failing compute-dy/dt (tapes_py_interface.scm:42-44, 81).
  Through its binding, which checks nothing, a failure here must not leave the zero-filled result
  buffer looking like dy/dt = 0: the result is NaN (all of it when the table size is known)."""
  import numpy
  os.environ['MARKOV_TAPES_QUIET'] = '1'
  u_lib = reference_binding()
  assert u_lib.c_register_problems(123) == 124
  u_lib.tapes_last_error.restype = ctypes.c_char_p
  # unknown tag: the size of the caller's buffer is not known, so the first entry carries the NaN
  out = reference_dy_dt(u_lib, 'no-such-problem', 3, numpy.full(8, 0.125))
  assert numpy.isnan(out[0])
  assert b'unknown problem tag' in u_lib.tapes_last_error()
  # known tag: either there is a device and the known answer comes back (markov_tapes.py:357-365),
  # or the call fails (no GPU in this container) and every entry is NaN
  out = reference_dy_dt(u_lib, '__canary_problem_radioactive_decay', 3, numpy.full(8, 0.125))
  if u_lib.tapes_last_error():
    assert numpy.isnan(out).all()
  else:
    assert out.tolist() == [0.375, 0.125, 0.125, -0.125, 0.125, -0.125, -0.125, -0.375]
  # a window length no table can have: still a NaN, still no write past the first entry
  out = reference_dy_dt(u_lib, '__canary_problem_radioactive_decay', 40, numpy.full(8, 0.125))
  assert numpy.isnan(out[0]) and (out[1:] == 0).all()
  capfd.readouterr()


def test_error_message_belongs_to_the_last_call():
  """c_compute_dy_dt clears the error on entry: a message left by an earlier unchecked call must
  not make a later successful call look failed (the mirror's get_dy_dt raises on any message)."""
  import numpy
  lib = _lib.load()
  lib.c_register_problems(0)
  bad = numpy.frombuffer(b'no-such-problem\x00', dtype=numpy.uint8)
  buf = numpy.zeros(8)
  lib.c_compute_dy_dt(bad.ctypes.data, 3, 0, buf.ctypes.data, buf.ctypes.data)
  assert lib.tapes_last_error()
  assert lib.tapes_rhs_device(None, None, None, None) == 1  # null handle: refused with a message
  assert b'null model' in lib.tapes_last_error()


def test_stale_and_foreign_handles_are_refused():
  """Handles outlive their structures (abi.cu ModelHandle): something that is not a handle is
  refused instead of dereferenced."""
  import numpy
  lib = _lib.load()
  junk = numpy.zeros(16, dtype=numpy.uint64)
  assert lib.tapes_sync(junk.ctypes.data) == 1
  assert b'not a model handle' in lib.tapes_last_error()
  lib.tapes_clear_error()
