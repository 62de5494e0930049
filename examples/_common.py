import os
import sys

os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
