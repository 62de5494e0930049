"""The reference's shipped small configurations: time per right-hand side, device-resident and
through the host-buffer entry point, with the multi-launch path (CUDA-graph replay) and with the
single-launch kernel at every cluster size; the CPU port (literal mode, as the reference evaluates)
beside it.  GPU box.  usage: time_small_r02.py"""
import os, sys, time
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt
from oracle import oracle

oracle.build()
CASES = [('ex2-ferromagnetic-chain', 2, 3, configs.ex2_p0(3, 1 / 250)), ('ex2-ferromagnetic-chain', 2, 7, configs.ex2_p0(7, 1 / 250)),
         ('ex3-copolymerization', 4, 6, configs.ex3_p0(6)), ('ex5-msrtf-machine', 5, 5, configs.ex5_p0(5)),
         ('ex4-chemical-turing', 9, 5, configs.ex4_p0(5, powered_fraction=0.04)),
         ('ex4var2-chemical-turing', 10, 5, configs.ex4var2_p0(5))]


def device_us(model, p, reps=2000):
  out = torch.empty_like(p)
  for _ in range(20):
    model.rhs(p, out)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    for _ in range(20):
      model.rhs(p, out)
    e0.record()
    for _ in range(reps):
      model.rhs(p, out)
    e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps * 1e3


def host_us(f, p0, reps=1000):
  for _ in range(20):
    f(p0, 0.0)
  t0 = time.perf_counter()
  for _ in range(reps):
    f(p0, 0.0)
  return (time.perf_counter() - t0) / reps * 1e6


for tag, a, k, p0 in CASES:
  model = device.DeviceModel(tag, k)
  f = mt.get_dy_dt(tag=tag, size_a=a, cl_k=k)
  # mid-trajectory table (full support where the dynamics fill it): what a stepper spends its time on
  mid = configs.markov_table(a, k, 3)
  p = torch.from_numpy(mid).cuda()
  t0 = time.perf_counter(); reps = 0
  while time.perf_counter() - t0 < 0.5:
    oracle.compute_dy_dt(tag, k, mid, mode=oracle.LITERAL); reps += 1
  cpu_us = (time.perf_counter() - t0) / reps * 1e6
  default_cluster = model.info['launches_per_rhs'] == 1
  model.set_option('fused_small', 0)
  line = [f'{tag} k={k} n={a ** k} nodes={model.info["n_nodes"]} nnz={model.info["nnz"]}: CPU port (literal) {cpu_us:.1f} us | '
          f'multi-launch ({model.info["launches_per_rhs"]} kernels, graph replay) device {device_us(model, p):.1f} us host {host_us(f, mid):.1f} us']
  model.set_option('fused_small', 1)
  for cluster in (1, 2, 4, 8, 16):
    model.set_option('fused_cluster', cluster)
    line.append(f'single launch x{cluster}: device {device_us(model, p):.1f} us host {host_us(f, mid):.1f} us')
  print(' | '.join(line), flush=True)
  mt.u_lib.tapes_release_model(tag.encode(), k)

# the reference's runs end to end
p0 = configs.ex4_p0(5, powered_fraction=0.04)
for fused in (0, 1):
  os.environ['TAPES_FUSED_SMALL'] = str(fused)
  mt.u_lib.tapes_release_model(b'ex4-chemical-turing', 5); mt.u_lib.tapes_release_model(b'ex5-msrtf-machine', 5)
  mt.model_stats(tag='ex4-chemical-turing', cl_k=5); mt.model_stats(tag='ex5-msrtf-machine', cl_k=5)
  t0 = time.perf_counter()
  _, st = mt.ode_integrate_device(tag='ex4-chemical-turing', size_a=9, cl_k=5, p0=p0, ts=numpy.linspace(0, 2000, 2001), rtol=1e-13,
                                  atol=1e-13, observables=[[0], [1], [6], [7], [5, 0], [5, 4, 1], [5, 4, 5, 2], [5, 4, 5, 4, 3]],
                                  return_states=False, want_stats=True)
  t1 = time.perf_counter()
  _, st5 = mt.ode_integrate_device(tag='ex5-msrtf-machine', size_a=5, cl_k=5, p0=configs.ex5_p0(5), ts=numpy.linspace(0, 500, 4001),
                                   rtol=1e-13, atol=1e-13, observables=[[0], [1], [2]], return_states=False, want_stats=True)
  t2 = time.perf_counter()
  print(f'single launch {fused}: ex4 DOP853 t=0..2000 1e-13, 8 observables x 2001 times: {t1 - t0:.3f} s ({st["nfev"]} rhs); '
        f'ex5 t=0..500, 4001 times: {t2 - t1:.3f} s ({st5["nfev"]} rhs)', flush=True)
