"""Turns an `ncu --set full` raw-page CSV of one right-hand side into (1) a table of selected columns
per launch (committed under profiles/) and (2) profiles/ncu_traffic.json: DRAM bytes per step and
kernel, which bench.py copies into the `traffic` keys when it runs the very same structure.

usage: ncu -i capture.ncu-rep --page raw --csv > raw.csv
       python scripts/ncu_traffic.py raw.csv profiles/r02_x_ncu_full_step.csv A k rules seed nnz n_nodes [--json]
"""
import csv, json, os, re, sys
from collections import OrderedDict

COLUMNS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
           'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'sm__maximum_warps_per_active_cycle_pct']


def scale_bytes(value, unit):
  return float(value.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[unit]


def scale_ms(value, unit):
  return float(value.replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}[unit]


def main():
  raw, out_csv = sys.argv[1], sys.argv[2]
  rows = list(csv.reader(open(raw)))
  head, units = rows[0], rows[1]
  name_i = head.index('Kernel Name')
  cols = [c for c in COLUMNS if c in head]
  kernels = OrderedDict()
  with open(out_csv, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['kernel'] + [f'{c} [{units[head.index(c)]}]' for c in cols])
    for r in rows[2:]:
      full = r[name_i]
      short = re.sub(r'^void ', '', full)
      short = re.sub(r'tapes::|<?unnamed>::|\(anonymous namespace\)::', '', short).split('(')[0].strip()
      w.writerow([short] + [r[head.index(c)] for c in cols])
      base = short.split('<')[0]
      k = kernels.setdefault(base, dict(dram_bytes=0.0, launches=0, ncu_ms=0.0))
      k['dram_bytes'] += scale_bytes(r[head.index('dram__bytes_read.sum')], units[head.index('dram__bytes_read.sum')])
      k['dram_bytes'] += scale_bytes(r[head.index('dram__bytes_write.sum')], units[head.index('dram__bytes_write.sum')])
      k['ncu_ms'] += scale_ms(r[head.index('gpu__time_duration.sum')], units[head.index('gpu__time_duration.sum')])
      k['launches'] += 1
  total_b = sum(k['dram_bytes'] for k in kernels.values())
  total_ms = sum(k['ncu_ms'] for k in kernels.values())
  for name, k in kernels.items():
    print(f'{name:28s} x{k["launches"]:3d} {k["ncu_ms"]:8.3f} ms {k["dram_bytes"] / 1e9:8.3f} GB '
          f'{k["dram_bytes"] / max(k["ncu_ms"], 1e-9) / 1e6:8.1f} GB/s')
  print(f'{"step":28s}      {total_ms:8.3f} ms {total_b / 1e9:8.3f} GB {total_b / total_ms / 1e6:8.1f} GB/s')
  if '--json' in sys.argv:
    a, k, rules, seed, nnz, n_nodes = [int(x) for x in sys.argv[3:9]]
    for v in kernels.values():
      v['source'] = out_csv
    rec = dict(source=f'{out_csv} (ncu --set full --clock-control none, every kernel of one right-hand side, B200; '
                      'python scripts/one_rhs.py; regenerate with scripts/ncu_traffic.py)',
               config=dict(size_a=a, cl_k=k, rules=rules, seed=seed, nnz=nnz, n_nodes=n_nodes), kernels=kernels)
    path = os.path.join(os.path.dirname(os.path.abspath(out_csv)), 'ncu_traffic.json')
    json.dump(rec, open(path, 'w'), indent=1)
    print('wrote', path)


if __name__ == '__main__':
  main()
