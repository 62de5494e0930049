"""Structure build of the bench configuration (or A k rules from argv), several times in one process:
the build timings the library records (kernels of the expansion without allocation, lists + CSR,
slices).  GPU box.  usage: time_build_r02.py [A k rules [repeats]]"""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, rules = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (10, 8, 24)
repeats = int(sys.argv[4]) if len(sys.argv) > 4 else 3
tag = f'build-{A}-{k}-{rules}'
mt.register_rule_set(tag, A, configs.random_rule_set(A, rules, seed=0))
for r in range(repeats):
  model = device.DeviceModel(tag, k)
  t, i = model.timing, model.info
  print(f"build {r}: expand kernels {t['device_expand_ms'] - t['expand_alloc_ms']:.1f} ms  (with allocation {t['device_expand_ms']:.1f})  "
        f"lists+csr {t['device_csr_ms']:.1f}  slices {t['device_slices_ms']:.1f}  nodes {i['n_nodes']}  inserts {i['hash_inserts']}  "
        f"unique {i['hash_unique']}  retries {i['hash_retries']}", flush=True)
  del model
  mt.u_lib.tapes_release_model(tag.encode(), k)
