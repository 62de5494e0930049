"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
name_i, val_i, unit_i = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
  if r[hdr.index('Metric Name')] != 'gpu__time_duration.sum':
    continue
  v = float(r[val_i].replace(',', ''))
  scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0, 'second': 1e3}[r[unit_i]]
  name = re.sub(r'\(.*', '', r[name_i]).replace('tapes::<unnamed>::', '').replace('unnamed>::', '')
  tot[name] += v * scale
  cnt[name] += 1
total = sum(tot.values())
print(f'{"kernel":48s} {"launches":>8s} {"ms":>10s} {"share":>7s}')
for name in sorted(tot, key=tot.get, reverse=True):
  print(f'{name[:48]:48s} {cnt[name]:8d} {tot[name]:10.3f} {tot[name] / total:7.1%}')
print(f'{"total":48s} {sum(cnt.values()):8d} {total:10.3f}')
