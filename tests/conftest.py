import json
import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
if GOLDEN not in sys.path:
  sys.path.insert(0, GOLDEN)  # the generating scripts hold the lists of observables the tests read


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def oracle():
  from oracle import oracle as orc
  orc.build()
  return orc


@pytest.fixture(scope='session')
def p0_fixtures():
  return numpy.load(os.path.join(GOLDEN, 'p0_fixtures.npz'))


@pytest.fixture(scope='session')
def trajectories():
  return numpy.load(os.path.join(GOLDEN, 'oracle_trajectories.npz'))


@pytest.fixture(scope='session')
def known_answers():
  with open(os.path.join(GOLDEN, 'reference_known_answers.json')) as f:
    return json.load(f)


def dense(idx, val, n):
  p = numpy.zeros(n, dtype=numpy.float64)
  p[idx] = val
  return p
