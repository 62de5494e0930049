"""Round-2 sweeps on the bench structure (A = 10, k = 8, 24 rules unless given): phases of one right-hand
side with the plane kernel / the left ratio table on and off, and the host-buffer entry point with
pinned and pageable buffers for several numbers of staging threads.
usage: sweep_r02.py [A k R]"""
import os, sys, time
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import bench
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 and sys.argv[1].isdigit() else (10, 8, 24)
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
p = bench.device_product_table(A, k, 3, torch.device('cuda'))
out = torch.empty_like(p)


def phases(model, reps=10):
  ph = numpy.zeros(3)
  for _ in range(3):
    model.rhs(p, out)
  for _ in range(reps):
    ph += model.rhs_profile(p, out)
  return ph / reps


want = None
for left in ('0', '1'):
  os.environ['TAPES_RATIO_LEFT'] = left
  mt.u_lib.tapes_release_model(tag.encode(), k)
  model = device.DeviceModel(tag, k)
  for plane in (1, 0):
    model.set_option('plane_kernel', plane)
    ph = phases(model)
    got = out.clone()
    if want is None:
      want = got
    print(f'ratio_left={left} plane_kernel={plane}: prepass {ph[0]:.3f} levels {ph[1]:.3f} product {ph[2]:.3f} total {ph.sum():.3f} ms '
          f'same_bits={bool(torch.equal(got, want))} plane_groups={model.info["plane_groups"]} launches={model.info["launches_per_rhs"]}', flush=True)
os.environ['TAPES_RATIO_LEFT'] = '0'
mt.u_lib.tapes_release_model(tag.encode(), k)
model = device.DeviceModel(tag, k)

if '--no-host' in sys.argv:
  sys.exit(0)
# host-buffer entry point
n = A ** k
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
h_in.copy_(p)
a_tag = numpy.frombuffer(tag.encode() + b'\x00', dtype=numpy.uint8)
y = h_in.numpy().copy()


def timed(call, reps=3):
  call(); call()
  t0 = time.perf_counter()
  for _ in range(reps):
    call()
  return 1e3 * (time.perf_counter() - t0) / reps


pinned = timed(lambda: mt.u_lib.c_compute_dy_dt(a_tag.ctypes.data, k, 0, h_in.data_ptr(), h_out.data_ptr()))
print(f'c_compute_dy_dt pinned in / pinned out: {pinned:.1f} ms', flush=True)
f = mt.get_dy_dt(tag=tag, size_a=A, cl_k=k)
for threads in ('4', '8', '12', '16'):
  os.environ['TAPES_COPY_THREADS'] = threads
  mt.u_lib.cleanup_gambit(None) if False else None
  res = numpy.empty(n)
  t_pageable = timed(lambda: mt.u_lib.c_compute_dy_dt(a_tag.ctypes.data, k, 0, y.ctypes.data, res.ctypes.data))
  t_half = timed(lambda: mt.u_lib.c_compute_dy_dt(a_tag.ctypes.data, k, 0, y.ctypes.data, h_out.data_ptr()))
  keep = [None]
  def drop_in():
    keep[0] = f(y, 0.0)
  t_drop = timed(drop_in)
  print(f'staging threads (first use decides): asked {threads}: pageable in / pageable out {t_pageable:.1f} ms, pageable in / pinned out '
        f'{t_half:.1f} ms, get_dy_dt(numpy) {t_drop:.1f} ms same_bits={bool(numpy.array_equal(res, h_out.numpy()))}', flush=True)
  break  # the pool is created once per process: one setting per run
