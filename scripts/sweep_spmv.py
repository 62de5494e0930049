"""Per-phase timings of one right-hand side for every SpMV lane setting (GPU box)."""
import os, sys
os.environ.setdefault('MARKOV_TAPES_QUIET', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import bench
from chemical_kinetics_and_program_execution_b200 import configs, device, markov_tapes as mt

A, k, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rules = configs.random_rule_set(A, R, seed=1)
tag = configs.synthetic_tag(A, R, 1)
mt.register_rule_set(tag, A, rules)
m = device.DeviceModel(tag, k)
n = A ** k
p = bench.device_product_table(A, k, 3, torch.device('cuda'))
out = torch.empty_like(p)
print(m.info, m.timing, flush=True)
ref = None
for lanes in (1, 2, 4, 8, 16):
    m.set_option('spmv_lanes', lanes)
    for _ in range(3): m.rhs(p, out)
    torch.cuda.synchronize()
    ph = numpy.zeros(3)
    for _ in range(5): ph += m.rhs_profile(p, out)
    ph /= 5
    o = out.cpu().numpy()
    if ref is None: ref = o
    dev = abs(o - ref).max() / abs(ref).max()
    gbs = bench.spmv_bytes(m.info['nnz'], n) / (ph[2] * 1e-3) / 1e9
    print(f'lanes={lanes:2d} phases_ms={ph} spmv={gbs:.0f} GB/s ({gbs/6555.2:.3f}) step={bench.step_bytes(m.info["nnz"], n, A)/(ph.sum()*1e-3)/1e9:.0f} GB/s dev_vs_lanes1={dev:.1e}', flush=True)
