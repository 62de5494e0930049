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
