"""Wall time of the reference's shipped small configurations through this library (GPU box)."""
import os, sys, time
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
from chemical_kinetics_and_program_execution_b200 import configs, markov_tapes as mt

def timed(label, fn):
  t0 = time.perf_counter(); out = fn(); dt = time.perf_counter() - t0
  print(f'{label}: {dt:.3f} s', out if isinstance(out, dict) else '', flush=True)

p0 = configs.ex4_p0(5, powered_fraction=0.04)
f = mt.get_dy_dt(tag='ex4-chemical-turing', size_a=9, cl_k=5)
f(p0, 0.0)
model = mt.u_lib.tapes_model(b'ex4-chemical-turing', 5)
for graphs in (0, 1):
  mt.u_lib.tapes_model_set(model, b'graphs', graphs)
  f(p0, 0.0); f(p0, 0.0)
  t0 = time.perf_counter()
  for _ in range(500): f(p0, 0.0)
  print(f'ex4 c_compute_dy_dt (host buffers), graphs={graphs}: {(time.perf_counter() - t0) / 500 * 1e6:.1f} us per call', flush=True)
  timed(f'ex4 device DOP853 t=0..200 rtol=atol=1e-13, no output points, graphs={graphs}', lambda: mt.ode_integrate_device(
      tag='ex4-chemical-turing', size_a=9, cl_k=5, p0=p0, ts=numpy.array([0.0, 200.0]), rtol=1e-13, atol=1e-13,
      observables=[[0]], return_states=False, want_stats=True)[1])
timed('ex4 device DOP853 t=0..2000 rtol=atol=1e-13, 8 observables', lambda: mt.ode_integrate_device(
    tag='ex4-chemical-turing', size_a=9, cl_k=5, p0=p0, ts=numpy.linspace(0, 2000, 2001), rtol=1e-13, atol=1e-13,
    observables=[[0], [1], [6], [7], [5, 0], [5, 4, 1], [5, 4, 5, 2], [5, 4, 5, 4, 3]], return_states=False,
    want_stats=True)[1])
timed('ex4 SciPy DOP853 through get_dy_dt t=0..200', lambda: mt.ode_integrate_ivp(
    tag='ex4-chemical-turing', size_a=9, cl_k=5, p0=p0, ts=numpy.linspace(0, 200, 201),
    ivp_kwargs=dict(method='DOP853', rtol=1e-13, atol=1e-13)).shape and None)
p5 = configs.ex5_p0(5)
timed('ex5 device DOP853 t=0..500 rtol=atol=1e-13', lambda: mt.ode_integrate_device(
    tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=p5, ts=numpy.linspace(0, 500, 4001), rtol=1e-13, atol=1e-13,
    observables=[[0], [1], [2]], return_states=False, want_stats=True)[1])
