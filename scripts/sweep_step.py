"""Per-phase timings of one right-hand side for the flux-structure formats and the level-kernel
load batching (GPU box).  usage: sweep_step.py A k R [formats...]   e.g. csr s32 s8 s4"""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import bench
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
formats = sys.argv[4:] or ['csr', 's32', 's16', 's8', 's4', 's2']
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
n = A ** k
p = bench.device_product_table(A, k, 3, torch.device('cuda'))
out = torch.empty_like(p)
PEAK = bench.measured_peak()[0]


def phases(m, reps=5):
  for _ in range(3):
    m.rhs(p, out)
  torch.cuda.synchronize()
  ph = numpy.zeros(3)
  for _ in range(reps):
    ph += m.rhs_profile(p, out)
  return ph / reps


ref = None
for fmt in formats:
  mt.u_lib.tapes_release_model(tag.encode(), k)
  if fmt == 'csr':
    os.environ['TAPES_FLUX_FORMAT'] = 'csr'
  else:  # sNN: slices with runs of at least NN lanes; nNN: the same without the fused right chain
    os.environ['TAPES_FLUX_FORMAT'] = 'slices'
    os.environ['TAPES_RUN_MIN_LANES'] = fmt[1:]
    os.environ['TAPES_LEVEL_FUSE'] = '0' if fmt[0] == 'n' else '1'

  m = device.DeviceModel(tag, k)
  if fmt != 'csr':
    for fu in (2, 3, 4, 6, 8):
      m.set_option('flux_unroll', fu)
      print(f'format={fmt} flux_unroll={fu} flux_ms={phases(m)[2]:.3f}', flush=True)
    m.set_option('flux_unroll', 4)
  ph = phases(m)
  o = out.cpu().numpy()
  if ref is None:
    ref = o
    print({k_: m.info[k_] for k_ in ('n_states', 'n_nodes', 'nnz', 'n_levels', 'hash_inserts', 'hash_unique')}, flush=True)
  dev = abs(o - ref).max() / abs(ref).max()
  nnz = m.info['nnz']
  gbs = bench.spmv_bytes(nnz, n) / (ph[2] * 1e-3) / 1e9
  facts = {k_: m.info[k_] for k_ in ('slice_words', 'runs', 'run_entries', 'column_entries', 'column_slots')}
  print(f'format={fmt:4s} flux_ms={ph[2]:.3f} algorithmic={gbs:.0f} GB/s ({gbs / PEAK:.3f}) levels_ms={ph[1]:.3f} '
        f'dev_vs_first={dev:.1e} words_per_nnz={facts["slice_words"] / max(nnz, 1):.3f} {facts} timing={m.timing}',
        flush=True)

for unroll in (1, 2, 4, 5, 8):
  m.set_option('level_unroll', unroll)
  ph = phases(m)
  o = out.cpu().numpy()
  print(f'level_unroll={unroll} levels_ms={ph[1]:.3f} flux_ms={ph[2]:.3f} marg_ms={ph[0]:.3f} '
        f'dev={abs(o - ref).max() / abs(ref).max():.1e} owned_parents={m.info["owned_parents"]} '
        f'deferred_groups={m.info["deferred_groups"]}', flush=True)
