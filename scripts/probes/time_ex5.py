"""Where the time of the ex5 device-resident run goes (GPU box)."""
import os, sys, time
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy
from chemical_kinetics_and_program_execution_b200 import configs, markov_tapes as mt
p5 = configs.ex5_p0(5)
kw = dict(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=p5, rtol=1e-13, atol=1e-13, observables=[[0], [1], [2]],
          return_states=False, want_stats=True)
t0 = time.perf_counter(); mt.model_stats(tag='ex5-msrtf-machine', cl_k=5); print(f'build {time.perf_counter() - t0:.3f} s')
for rep in range(3):
  for n_out in (2, 4001):
    t0 = time.perf_counter()
    st = mt.ode_integrate_device(ts=numpy.linspace(0, 500, n_out), **kw)[1]
    print(f'rep {rep} outputs {n_out}: {time.perf_counter() - t0:.3f} s {st}', flush=True)
model = mt.u_lib.tapes_model(b'ex5-msrtf-machine', 5)
mt.u_lib.tapes_model_set(model, b'graphs', 0)
t0 = time.perf_counter()
st = mt.ode_integrate_device(ts=numpy.linspace(0, 500, 2), **kw)[1]
print(f'graphs off, outputs 2: {time.perf_counter() - t0:.3f} s', flush=True)
