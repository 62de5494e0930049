"""Product kernel vs the minimum lanes of a run in the sliced flux structure (bench structure).
usage: sweep_run_lanes.py [lanes ...]"""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import bench
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = 10, 8, 24
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
p = bench.device_product_table(A, k, 3, torch.device('cuda'))
out = torch.empty_like(p)
want = None
for lanes in [int(x) for x in sys.argv[1:]] or [32, 16, 10, 8]:
  os.environ['TAPES_RUN_MIN_LANES'] = str(lanes)
  mt.u_lib.tapes_release_model(tag.encode(), k)
  model = device.DeviceModel(tag, k)
  for unroll in (2, 3, 4, 6):
    model.set_option('flux_unroll', unroll)
    for _ in range(3):
      model.rhs(p, out)
    ph = numpy.zeros(3)
    for _ in range(10):
      ph += model.rhs_profile(p, out)
    ph /= 10
    if want is None:
      want = out.clone()
    err = float((out - want).abs().max() / want.abs().max())
    i = model.info
    print(f'min_run_lanes={lanes} flux_unroll={unroll}: product {ph[2]:.3f} ms (levels {ph[1]:.3f}) words={i["slice_words"]} runs={i["runs"]} '
          f'run_entries={i["run_entries"]} column_entries={i["column_entries"]} column_slots={i["column_slots"]} '
          f'slices_build_ms={model.timing["device_slices_ms"]:.1f} max_rel_dev={err:.1e}', flush=True)
