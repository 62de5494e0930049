"""B200-native tape-multiverse probability evolution.

`markov_tapes` is the drop-in module (same API as the reference's framework/markov_tapes.py);
importing it initialises the CUDA runtime and runs the reference's load-time known-answer test,
so it needs a B200.  `configs` (initial distributions, rule-set generators) and `_lib` (the
ctypes binding) import without a GPU.
"""
