"""Builds tapes_py_interface.so in-tree with nvcc for sm_100a.

The library name and location match what the reference's Python layer loads
(framework/markov_tapes.py:40-41): `tapes_py_interface.so` next to `markov_tapes.py`.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'tapes_py_interface.so')
SOURCES = ['abi.cu', 'engine.cu', 'flux.cu', 'integrate.cu', 'validate.cu', 'montecarlo.cu', 'hostcopy.cu', 'rules.cpp',
           'problems.cpp']
HEADERS = ['engine.h', 'rules.h', 'primitives.cuh', 'integrate.h', 'validate.h', 'montecarlo.h', 'cuda_check.h', 'hostcopy.h',
           os.path.join('..', '..', 'include', 'tapes_b200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC', '-shared',
    '-Xptxas', '-v',
]


def _nvcc():
  for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError('nvcc not found')


def needs_build():
  if not os.path.exists(OUT):
    return True
  built = os.path.getmtime(OUT)
  paths = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
  return any(os.path.exists(p) and os.path.getmtime(p) > built for p in paths)


def build(force=False, verbose=False):
  if not force and not needs_build():
    return OUT
  srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
  cmd = [_nvcc()] + NVCC_FLAGS + ['-o', OUT] + srcs
  proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  log = os.path.join(HERE, 'build.log')
  with open(log, 'w') as f:
    f.write(' '.join(cmd) + '\n' + proc.stdout)
  if verbose or proc.returncode != 0:
    sys.stderr.write(proc.stdout)
  if proc.returncode != 0:
    raise RuntimeError(f'nvcc failed (see {log})')
  return OUT


if __name__ == '__main__':
  build(force='--force' in sys.argv, verbose=True)
