"""SASS of selected kernels of the built library: mnemonic histogram and full listing.
usage: python scripts/sass_listing.py OUT.txt 'regex of demangled kernel names' [more regexes]"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'chemical_kinetics_and_program_execution_b200', 'tapes_py_interface.so')


def main():
  out_path, patterns = sys.argv[1], [re.compile(p) for p in sys.argv[2:]]
  sass = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True, check=True).stdout
  blocks = re.split(r'\n\s*Function : ', sass)
  with open(out_path, 'w') as f:
    f.write(f'# cuobjdump -sass {os.path.relpath(SO, ROOT)} (sm_100a), kernels matching {[p.pattern for p in patterns]}\n')
    for b in blocks[1:]:
      mangled = b.split('\n', 1)[0].strip()
      name = subprocess.run(['c++filt', mangled], capture_output=True, text=True).stdout.strip()
      if not any(p.search(name) for p in patterns):
        continue
      ops = collections.Counter()
      for line in b.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
        if m:
          ops[m.group(1)] += 1
      f.write(f'\n===== {name}\n')
      f.write(f'instructions: {sum(ops.values())}\n')
      groups = collections.Counter()
      for op, c in ops.items():
        groups[op.split('.')[0]] += c
      f.write('by mnemonic: ' + ', '.join(f'{op} {c}' for op, c in groups.most_common()) + '\n')
      f.write('memory and special: ' + ', '.join(f'{op} {c}' for op, c in sorted(ops.items())
                                                  if re.match(r'LD|ST|ATOM|RED|MUFU|UBLKCP|UTMA|LDGSTS|SYNCS|BAR|SHFL|DMUL|DFMA|DADD|DSETP', op)) + '\n')
      f.write(b)
  print('wrote', out_path)


if __name__ == '__main__':
  main()
