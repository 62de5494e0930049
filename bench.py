"""Benchmark of the master-equation step  dy/dt = S * w(p)  on a synthetic random rewrite-rule
multiverse (BASELINE.json: last config).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the CPU port of the reference path

One "step" is one full right-hand side over the whole state table: marginal tables, leaf-world
probabilities, every level of the window-extension forest (the p-dependent re-evaluation of the
rate structure) and the CSR product S*w.  With N > 1 every rank gets its own 24 rules (weak
scaling) and a step also contains the flux exchange, by default fused into the product kernel over
NVLink peer memory (parallel.PeerExchangeRhs).

`value` is the survey's metric: CSR-EQUIVALENT algorithmic GB/s of the
whole step, i.e. the bytes a plain CSR SpMV with per-step rate re-evaluation would move,
  step_bytes = 28 * nnz + 24 * n + 8 * n * (1 + 2 / (A - 1)),
divided by the step time.  It is a work rate, not a bandwidth: the kernels stream a compressed
structure (implicit children, runs of 32 states), so it can exceed the HBM peak.  The bandwidth
figures are `value_dram_gbs` / `dram_frac` (whole step: DRAM bytes of every kernel of a step from
the committed ncu capture of this very structure, profiles/ncu_traffic.json, over the step time)
and the `roofline*` objects, whose `achieved` counts the bytes the shipped format must move at the
least (every array a kernel streams read once, every result written once: DESIGN.md section 4),
which is below the real traffic, so `frac` <= `dram_frac` <= 1.  `roofline` is the product kernel
(flux_slices_kernel, the longest single launch), `roofline_levels` the level_kernel launches summed.
`parity_check`: the dy/dt of the TIMED workload is checked inside this run (sum over states = 0, and
against the CPU port evaluated on the same rules and table in the cpu_baseline leg).
Prints exactly one JSON line on rank 0.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

os.environ.setdefault('MARKOV_TAPES_QUIET', '1')  # stdout carries exactly one JSON line

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'master-eq SpMV GB/s (whole dy/dt step incl. rate re-evaluation; CSR-equivalent algorithmic bytes)'
UNIT = 'GB/s'


# stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent
# to stderr by pointing file descriptor 1 at it; the line itself goes to the original descriptor.
# Done by main() only, so that scripts importing this module keep their stdout.
_REAL_STDOUT = None


def claim_stdout():
  global _REAL_STDOUT
  if _REAL_STDOUT is None:
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
  sys.stdout.flush()
  os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + '\n').encode())


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--size-a', type=int, default=10)
  ap.add_argument('--cl-k', type=int, default=8)
  ap.add_argument('--rules-per-gpu', type=int, default=24)
  ap.add_argument('--seed', type=int, default=1)
  ap.add_argument('--chunks', type=int, default=4,
                  help='rounds of the peer exchange / row blocks of the overlapped NCCL exchanges (0 or 1 = no overlap)')
  ap.add_argument('--exchange', default='peer', choices=['peer', 'rs_ag', 'allreduce'],
                  help='flux exchange for N > 1: fused into the product over NVLink peer memory, '
                       'NCCL reduce-scatter + all-gather, or NCCL all-reduce in row blocks')
  ap.add_argument('--e2e-steps', type=int, default=3)
  ap.add_argument('--cpu-rules', type=int, default=0,
                  help='rules in the CPU-baseline sample (0 = one per usable host core, at most all)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-prewarm', action='store_true', help='do not touch the device memory once before the structure build')
  ap.add_argument('--no-strong-leg', action='store_true', help='N > 1, weak scaling: skip the strong-scaling measurement the line carries')
  ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                  help='N > 1: weak = rules-per-gpu rules on every rank (work grows with N), strong = the '
                       'N = 1 problem (rules-per-gpu rules in all) dealt to the N ranks')
  ap.add_argument('--ref-budget-s', type=float, default=600.0,
                  help='--impl reference: seconds of CPU work the timed steps may take in all (fewer steps '
                       'are timed, and reported, when `steps` full-table evaluations do not fit)')
  return ap.parse_args()


def step_bytes(nnz, n, size_a):
  return 28.0 * nnz + 24.0 * n + 8.0 * n * (1.0 + 2.0 / max(size_a - 1, 1))


def spmv_bytes(nnz, n):
  return 12.0 * nnz + 16.0 * n


def level_bytes(info):
  """Compulsory bytes of the level evaluation of one step (plane_kernel + level_kernel launches and
  prefix_sums_kernel): what one launch must move if every byte it touches came
  from DRAM exactly once.  Per level the table / ratio entries its nodes read, each entry once however
  many nodes (of different seeds) read it: min(nodes of the level, states) * 8 B, summed over the levels
  (model info `distinct_table_reads`); one 8-byte weight write per node that is written (the right
  children of groups in regular blocks are not, unless a later level reads them); 25 B per left-parent
  record (ids, length, parent weight, short marginal); per prefix group 16 B of records outside regular
  blocks (48 B per 256 groups inside), its sum written (8 B), read by the level that evaluates its
  children (8 B) and by prefix_sums_kernel with its number (12 B); 8 B per parent weight that is
  gathered (a node is the parent of one group); 16 B per prefix for the per-prefix sums.  The real
  traffic is above this by the L2 misses of repeated reads."""
  nodes, size_a = info['n_nodes'], max(info['alphabet'], 1)
  gathered = info['hash_inserts'] - info.get('owned_parents', 0)
  plane = info.get('plane_groups', 0)  # 48 bytes per block of 256 such groups instead of 16 per group
  groups, prefixes = info['hash_unique'], info['n_states'] // size_a
  unwritten = 0 if info.get('materialize_right', 1) else plane * size_a
  per_group = 16.0 * (groups - plane) + 48.0 * plane / 256
  sums = 8.0 * groups + 8.0 * info.get('deferred_groups', 0) + 12.0 * groups + 16.0 * prefixes
  table_reads = 8.0 * info.get('distinct_table_reads', nodes)
  return table_reads + 8.0 * (nodes - unwritten) + 25.0 * info.get('left_parents', 0) + per_group + 8.0 * gathered + sums


def flux_format_bytes(info, n):
  """Compulsory bytes of the product kernel in the shipped format (sliced flux structure of the stored
  entries + per-group flux of the right children, csrc/flux.cu, csrc/flux_device.cuh), every byte
  counted once: slice pointers and run counts, every structure word (column padding included: the
  kernel has to read it), the weight of every stored term once (a term is read by its source and its
  destination row; the second read is credited to L2); per prefix group its inflow list entry (8 B)
  and its sum (8 B); the ratio table once (8 B per state: the outflow reads a state's own ratio, the
  inflows read ratios of other prefixes - repeats credited to L2); per prefix the list offset and the
  outflow sum (16 B); the result (8 B per state).  The real traffic is above this by the L2 misses
  of the repeated reads."""
  slices, size_a = info.get('n_slices', (n + 31) // 32), max(info.get('alphabet', 1), 1)
  grouped = info.get('nnz_stored', info['nnz']) < info['nnz']
  groups = info.get('hash_unique', 0) if grouped else 0
  right_children = groups * size_a
  lists = 16.0 * groups + (8.0 * n + 16.0 * (n // size_a) if grouped else 0.0)
  return (8.0 * (slices + 1) + 4.0 * slices + 4.0 * info.get('slice_words', 0) + 8.0 * (info['n_terms'] - right_children)
          + lists + 8.0 * n)


def prepass_bytes(info, n, size_a):
  """Marginal tables and per-step ratio tables.  With the right table only (the default) the longest
  marginal table and the ratios come from one pass over the table (marginal_ratio_kernel): the table
  read once, one ratio per state written, the shorter marginal tables written and read once each.
  With the left table as well: the table read once more and a second ratio per state written."""
  shorter = 16.0 * n / max(size_a - 1, 1)
  if info.get('ratio_tables', 1) <= 1:
    return 16.0 * n + shorter
  return 8.0 * n + shorter + 8.0 * n + 16.0 * n


_TRAFFIC = None


def recorded_traffic(kernel, info, args):
  """DRAM bytes per step of `kernel` (all its launches of one right-hand side summed) from the
  committed ncu capture, when that capture was taken on this very structure; else None.
  kernel = None: the whole step (every kernel of the capture)."""
  global _TRAFFIC
  try:
    if _TRAFFIC is None:
      with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
        _TRAFFIC = json.load(f)
    rec = _TRAFFIC
    same = all(rec['config'].get(k) == v for k, v in
               dict(size_a=args.size_a, cl_k=args.cl_k, rules=args.rules_per_gpu, seed=args.seed,
                    nnz=info['nnz'], n_nodes=info['n_nodes']).items())
    if not same:
      return None
    if kernel is None:
      return float(sum(k['dram_bytes'] for k in rec['kernels'].values()))
    return float(rec['kernels'][kernel]['dram_bytes'])
  except Exception:
    return None


def workload_config(args, world):
  """`config` of the JSON line: names the workload only, so that this repo's arm and the reference
  arm print the same object for the same command line."""
  strong = getattr(args, 'scaling', 'weak') == 'strong'
  total_rules = args.rules_per_gpu if strong else args.rules_per_gpu * world
  return dict(workload='synthetic-random-rewrite-rules', size_a=args.size_a, cl_k=args.cl_k,
              n_states=args.size_a ** args.cl_k, rules_per_gpu=args.rules_per_gpu, total_rules=total_rules,
              seed=args.seed, table='full-support product table (Dirichlet symbol frequencies, seed + 2)',
              parallelism=('single GPU' if world == 1 else
                           f'{total_rules} rules dealt to {world} ranks, one flux exchange per step'),
              l2='inputs larger than L2 (table, weights and flux structure each exceed 126 MB)')


def usable_cpu_workers(n_states, wanted):
  """Worker processes for the CPU port: one per rule and core, bounded by host memory (each worker
  holds the table, its marginals, the result and the accumulators: about 5 tables)."""
  cores = len(os.sched_getaffinity(0))
  try:
    import psutil
    by_mem = int(psutil.virtual_memory().available * 0.6 // (5 * 8 * n_states + (1 << 28)))
  except Exception:
    by_mem = 4
  return max(1, min(cores, wanted, by_mem))


def measured_peak():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  except Exception:
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
  """Samples SM clocks and throttle reasons through NVML every 20 ms while the timed region runs
  (the same counters as `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`)."""

  def __init__(self, index):
    self.index = index
    self.samples = []
    self.stop = threading.Event()
    self.thread = threading.Thread(target=self._run, daemon=True)
    self.max_mhz = None
    self.error = None

  def _run(self):
    try:
      import pynvml
      pynvml.nvmlInit()
      visible = os.environ.get('CUDA_VISIBLE_DEVICES')
      phys = int(visible.split(',')[self.index]) if visible and visible.split(',')[self.index].isdigit() else self.index
      h = pynvml.nvmlDeviceGetHandleByIndex(phys)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
      get_reasons = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
          pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
      while not self.stop.is_set():
        self.samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), int(get_reasons(h))))
        self.stop.wait(0.02)
    except Exception as ex:  # NVML binding unavailable: fall back to polling nvidia-smi
      self.error = repr(ex)
      self._run_smi()

  def _run_smi(self):
    query = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
    bits = [0x8, 0x40, 0x20, 0x4]
    while not self.stop.is_set():
      try:
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={query}',
                              '--format=csv,noheader,nounits'], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(',')]
        if len(f) >= 6 and f[0].isdigit():
          self.max_mhz = int(f[1])
          mask = sum(b for b, v in zip(bits, f[2:6]) if v.lower().startswith('active'))
          self.samples.append((int(f[0]), mask))
      except Exception:
        pass
      self.stop.wait(0.1)

  def __enter__(self):
    self.thread.start()
    time.sleep(0.05)
    return self

  def __exit__(self, *exc):
    self.stop.set()
    self.thread.join(timeout=10)

  def summary(self):
    # NVML clocks-event-reason bits
    bits = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}
    sm = sorted(s[0] for s in self.samples)
    seen = 0
    for s in self.samples:
      seen |= s[1]
    out = dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=self.max_mhz,
               reasons=sorted(name for bit, name in bits.items() if seen & bit), samples=len(self.samples))
    if self.error:
      out['error'] = self.error
    return out


def make_workload(args, world, rank):
  """The synthetic rule set and this rank's share.  Weak scaling with exactly equal work per GPU:
  the base set of `rules_per_gpu` random rules is the N = 1 workload; at N GPUs the problem is the
  union of N copies of it with the alphabet rotated by the rank (different rules, identical
  structure sizes), and rank g evaluates copy g."""
  from chemical_kinetics_and_program_execution_b200 import configs, parallel
  base = configs.random_rule_set(args.size_a, args.rules_per_gpu, seed=args.seed)
  r = args.rules_per_gpu
  if getattr(args, 'scaling', 'weak') == 'strong':
    # strong scaling: the N = 1 problem itself, its rules dealt to the ranks by their term counts
    rules = base
    local = parallel.split_rule_set(rules, world, rank, args.size_a, args.cl_k) if world > 1 else rules
    tag = configs.synthetic_tag(args.size_a, r, args.seed) + (f'-part{rank}of{world}' if world > 1 else '')
    return rules, local, tag
  rules = configs.concat_rule_sets([configs.rotated_rule_set(base, g, args.size_a) for g in range(world)])
  local = parallel.take_rules(rules, numpy.arange(rank * r, (rank + 1) * r)) if world > 1 else rules
  tag = configs.synthetic_tag(args.size_a, r * world, args.seed) + (f'-rank{rank}of{world}' if world > 1 else '')
  return rules, local, tag


def device_product_table(size_a, cl_k, seed, device):
  """Full-support product table built on the device (Dirichlet symbol frequencies)."""
  import torch
  rng = numpy.random.default_rng(seed)
  f = torch.from_numpy(rng.dirichlet(numpy.ones(size_a))).to(device)
  table = f
  for _ in range(cl_k - 1):
    table = (table[:, None] * f[None, :]).reshape(-1)
  return table.contiguous()


# --------------------------------------------------------------------------------------------
# CPU port (oracle) legs: the only places bench.py executes anything under oracle/.
# --------------------------------------------------------------------------------------------
_SHARED_ROWS = None  # anonymous shared mapping the forked workers write their dy/dt into


def _cpu_worker(job):
  tag, size_a, cl_k, rules, seed, row = job
  from chemical_kinetics_and_program_execution_b200 import configs
  from oracle import oracle
  oracle.register_rules(tag, size_a, rules)
  f = numpy.random.default_rng(seed).dirichlet(numpy.ones(size_a))
  p = configs.product_table(f, cl_k)
  t0 = time.perf_counter()
  out, counters = oracle.compute_dy_dt(tag, cl_k, p, mode=oracle.MERGED, want_counters=True)
  seconds = time.perf_counter() - t0
  if _SHARED_ROWS is not None and row is not None:  # not timed: hands the result to the parent
    n = out.size
    numpy.frombuffer(_SHARED_ROWS, dtype=numpy.float64, count=n, offset=8 * n * row)[:] = out
  return seconds, counters


def cpu_port_step(args, rules_subset, reps=1, want_result=False):
  """Times the merged-mode CPU port on `rules_subset` (as a problem of its own: the selection
  weights of the subset), one worker process per share of the rules (the forests of different
  rules are independent, which is the only parallelism the path offers on a CPU).  Returns
  (seconds, workers, counters, dy/dt or None); with want_result the workers leave their partial
  dy/dt in shared memory and the parent adds them up in worker order."""
  global _SHARED_ROWS
  import mmap
  import multiprocessing as mp
  from chemical_kinetics_and_program_execution_b200 import parallel
  n_rules = len(rules_subset['rate'])
  cl_k = getattr(args, 'cpu_cl_k', None) or args.cl_k
  n = args.size_a ** cl_k
  cores = usable_cpu_workers(n, n_rules)
  jobs = [(f'cpu-sample-{i}', args.size_a, cl_k, parallel.split_rule_set(rules_subset, cores, i), args.seed + 2,
           i if want_result else None) for i in range(cores)]
  _SHARED_ROWS = mmap.mmap(-1, 8 * n * cores) if want_result else None  # inherited by the forked workers
  ctx = mp.get_context('fork')
  times, counters, total = [], None, None
  try:
    for _ in range(reps):
      with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
      times.append(max(r[0] for r in res))  # workers run concurrently; table set-up is not timed
      counters = {k: sum(r[1][k] for r in res) for k in res[0][1]}
    if want_result:
      rows = numpy.frombuffer(_SHARED_ROWS, dtype=numpy.float64, count=n * cores).reshape(cores, n)
      total = rows[0].copy()
      for i in range(1, cores):
        total += rows[i]
      del rows
  finally:
    if _SHARED_ROWS is not None:
      try:
        _SHARED_ROWS.close()
      except BufferError:
        pass
      _SHARED_ROWS = None
  return min(times), cores, counters, total


def _guarded(fn, *fn_args):
  """An auxiliary figure must not cost the bench line."""
  try:
    return fn(*fn_args)
  except Exception as ex:  # pylint: disable=broad-except
    return dict(error=repr(ex))


def literal_vs_merged(args, rules, n_rules=4):
  """beside the states this library expands (every distinct node once), the
  number the reference's literal recursion visits for the same rules (right-shifted windows are
  revisited once per left context, about A^(2k-3) per one-cell rule).  Counted exactly by the CPU port
  in both modes at the window lengths where the literal recursion finishes in a second, then
  extrapolated to the bench's cl_k with the last measured growth factor per unit of k."""
  from oracle import oracle
  from chemical_kinetics_and_program_execution_b200 import configs
  sample = {k: numpy.asarray(v)[:n_rules] for k, v in rules.items()}
  oracle.register_rules('bench-literal-count', args.size_a, sample)
  rows = []
  for k in range(2, args.cl_k + 1):
    if float(args.size_a) ** (2 * k - 3) * n_rules > 6e7:
      break
    p = configs.dirichlet_product_table(args.size_a, k, 3)
    lit = oracle.compute_dy_dt('bench-literal-count', k, p, mode=oracle.LITERAL, want_counters=True)[1]
    mer = oracle.compute_dy_dt('bench-literal-count', k, p, mode=oracle.MERGED, want_counters=True)[1]
    rows.append(dict(cl_k=k, literal_nodes=lit['ext_nodes'] + lit['worlds'], literal_accumulate_calls=lit['acc_calls'],
                     merged_nodes=mer['ext_nodes'] + mer['worlds'], merged_terms=mer['acc_calls']))
  out = dict(rules=n_rules, counted=rows)
  if len(rows) >= 2 and rows[-1]['cl_k'] < args.cl_k:
    steps = args.cl_k - rows[-1]['cl_k']
    g_lit = rows[-1]['literal_nodes'] / max(rows[-2]['literal_nodes'], 1)
    g_mer = rows[-1]['merged_nodes'] / max(rows[-2]['merged_nodes'], 1)
    out['extrapolated'] = dict(cl_k=args.cl_k, literal_nodes=rows[-1]['literal_nodes'] * g_lit ** steps,
                               merged_nodes=rows[-1]['merged_nodes'] * g_mer ** steps,
                               growth_per_k=dict(literal=g_lit, merged=g_mer),
                               note='the reference visits literal_nodes for what is merged_nodes distinct nodes here; '
                                    'the literal growth factor is still rising towards A^2 at the last counted '
                                    'cl_k, so the extrapolation understates the reference\'s work')
  return out


def run_reference(args):
  """--impl reference: the CPU port of the reference's compute-dy/dt (oracle, merged mode; the
  Gambit-C original cannot be built in this image) on the bench's own configuration: the same
  table (size_a ** cl_k states) and rank 0's rules_per_gpu rules, one worker process per share of
  the rules on all usable host cores.  At N > 1 (weak scaling) that is a bounded sample of the job -
  one rank's rules of the N * rules_per_gpu - at the same per-rule work, so the GB/s figure is the
  one the whole job would show.  warmup + steps full-table evaluations are run when they fit
  --ref-budget-s seconds of CPU work; when they do not, the warm-up goes first (the CPU path has
  nothing to warm up: no cached structures, no clocks to ramp - every evaluation rebuilds the tree
  like the reference, tm.scm:1476), then timed steps; `warmup` and `steps` of the line say what was
  actually run."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  from chemical_kinetics_and_program_execution_b200 import configs
  from oracle import oracle
  oracle.build()
  world = max(1, args.gpus)
  config = workload_config(args, world)
  rules = configs.random_rule_set(args.size_a, args.rules_per_gpu, seed=args.seed)  # rank 0's rules
  n = args.size_a ** args.cl_k
  t1, cores, counters, _ = cpu_port_step(args, rules)
  steps, warmup = max(1, args.steps), max(0, args.warmup)
  if (warmup + steps) * t1 > args.ref_budget_s:
    warmup = 0
    steps = max(1, min(steps, int(args.ref_budget_s // max(t1, 1e-9))))
  every = [t1]
  while len(every) < warmup + steps:
    every.append(cpu_port_step(args, rules)[0])
  times = every[warmup:]
  nnz = 2 * counters['acc_calls']
  bytes_step = step_bytes(nnz, n, args.size_a)
  ms = 1e3 * sum(times) / len(times)
  value = bytes_step / (ms * 1e-3) / 1e9
  sample_desc = (f'{args.rules_per_gpu} of {config["total_rules"]} rules (rank 0\'s share) on the full {n}-state table '
                 f'(cl_k={args.cl_k}, the bench configuration), nnz={nnz}, merged-mode CPU port of compute-dy/dt, '
                 f'{cores} worker processes, {warmup} warm-up + {len(times)} timed steps (requested {args.warmup} + {args.steps})')
  line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus,
              steps=len(times), warmup=warmup, ms_per_step=ms, higher_is_better=True,
              scaling=args.scaling, vs_baseline=None, dtype='f64', data='synthetic',
              config=config, same_config=True, steps_requested=args.steps, warmup_requested=args.warmup,
              cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind='port', sample=sample_desc),
              e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
              states_expanded_per_s=(counters['ext_nodes'] + counters['worlds']) / (ms * 1e-3),
              gpu_launches=0)
  emit(line)


# --------------------------------------------------------------------------------------------
# CUDA path
# --------------------------------------------------------------------------------------------
def prewarm_device_memory(device, fraction=0.45):
  """Warm-up of the driver's allocator, like the warm-up steps of the timed loop: the first process
  that maps a large part of a fresh box's HBM pays the driver 0.7-3 s for it (seen as
  build.expand_alloc_ms of 700-3000 ms in the first bench of a box and 20-30 ms in the second,
  profiles/README.md), which says nothing about the structure build.  One allocation of `fraction` of
  the free memory is written once and handed back before the build is timed; what it cost is reported
  in build.prewarm."""
  import torch
  t0 = time.perf_counter()
  free, _total = torch.cuda.mem_get_info(device)
  n_bytes = int(free * fraction)
  try:
    block = torch.empty(n_bytes, dtype=torch.uint8, device=device)
    block.zero_()
    torch.cuda.synchronize(device)
    del block
    torch.cuda.empty_cache()
  except RuntimeError as ex:  # out of memory on a shared device: the build will say so itself
    return dict(bytes=0, seconds=time.perf_counter() - t0, error=str(ex)[:120])
  return dict(bytes=n_bytes, seconds=time.perf_counter() - t0)


def run_b200(args):
  import torch
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit('launch with torch.distributed.run for --gpus > 1')
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    # the exchange runs beside the memory-bound product kernel: give NCCL's stream priority so its
    # few blocks are placed as soon as slots free up instead of queueing behind the product's grid
    os.environ.setdefault('TORCH_NCCL_HIGH_PRIORITY', '1')
    dist.init_process_group('nccl', device_id=device)

  from chemical_kinetics_and_program_execution_b200 import device as dev, markov_tapes as mt, parallel

  n = args.size_a ** args.cl_k
  rules, local_rules, tag = make_workload(args, world, rank)
  mt.register_rule_set(tag, args.size_a, local_rules)
  prewarm = None if args.no_prewarm else prewarm_device_memory(device)
  t0 = time.perf_counter()
  model = dev.DeviceModel(tag, args.cl_k)
  torch.cuda.synchronize()
  build_s = time.perf_counter() - t0
  info, timing = model.info, model.timing

  p = device_product_table(args.size_a, args.cl_k, args.seed + 2, device)
  out = torch.empty_like(p)
  sharded = None
  if world > 1:
    if args.exchange == 'peer':
      try:
        sharded = parallel.PeerExchangeRhs(model, rounds=max(args.chunks, 1) if args.chunks <= 16 else 16)
        padded = n
      except RuntimeError as ex:  # raised on all ranks together: no CUDA IPC between these processes
        if rank == 0:
          print(f'peer exchange unavailable ({ex}); using the NCCL all-reduce exchange', file=sys.stderr)
        args.exchange = 'allreduce'
    if args.exchange == 'peer':
      pass
    elif args.exchange == 'allreduce':
      sharded = parallel.OverlappedAllReduceRhs(model.weights, model.flux_rows, n, chunks=max(args.chunks, 1))
      padded = n
    elif args.chunks > 0:
      sharded = parallel.OverlappedRhs(model.weights, model.flux_rows, n, chunks=args.chunks, device=device)
      padded = sharded.padded
    else:
      sharded = parallel.ShardedRhs(lambda pin, pout: model.rhs(pin, pout), n, device=device)
      padded = sharded.padded
    p_full = torch.zeros(padded, dtype=torch.float64, device=device)
    p_full[:n] = p
    out_full = torch.zeros_like(p_full)

  def one_step(table=None):
    if sharded is None:
      model.rhs(p if table is None else table, out)
    elif args.exchange == 'peer':
      sharded.rhs_full(p_full if table is None else table)
    else:
      sharded.rhs_full(p_full if table is None else table, out_full)

  for _ in range(args.warmup):
    one_step()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  with ClockSampler(local_rank) as clocks:
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
      one_step()
    ev1.record()
    torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  ms_total = ev0.elapsed_time(ev1)
  if world > 1:
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
  ms_step = ms_total / args.steps

  # per-rank compute time of one step and the bare exchange, to explain the scaling
  rank_ms, comm_ms = None, None
  if world > 1:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
      model.rhs(p_full[:n], out_full[:n])
    e1.record()
    torch.cuda.synchronize()
    mine = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=device)
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    rank_ms = [float(x.item()) for x in every]
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
      if args.exchange == 'peer':  # no separate exchange step exists: it runs inside / beside the product
        pass
      elif args.exchange == 'allreduce':
        dist.all_reduce(out_full, op=dist.ReduceOp.SUM)
      else:
        dist.reduce_scatter_tensor(sharded.mine if hasattr(sharded, 'mine') else sharded.owned.view(-1)[:sharded.padded // world],
                                   out_full, op=dist.ReduceOp.SUM)
        dist.all_gather_into_tensor(out_full, sharded.mine if hasattr(sharded, 'mine') else sharded.owned.view(-1)[:sharded.padded // world])
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    comm_ms = float(t.item()) if args.exchange != 'peer' else None
    if args.exchange == 'peer':
      sharded.check()

  # whole-job structural size
  sizes = torch.tensor([info['nnz'], info['n_nodes'], info['n_terms'], info['launches_per_rhs'],
                        info['n_nodes'] + info['worlds_walked']], dtype=torch.float64, device=device)
  # expansion time with and without the part spent inside cudaMalloc / cudaFree: mapping ~50 GB of
  # fresh device memory costs 0.04-1.4 s depending on the state the box's driver is in (same build,
  # same binary: profiles/r01_n_build_cold_vs_warm.log), which says nothing about the expansion kernels
  expand_s = torch.tensor([(timing['device_expand_ms'] + timing['host_enumerate_ms']) * 1e-3,
                           (timing['device_expand_ms'] - timing['expand_alloc_ms'] + timing['host_enumerate_ms']) * 1e-3],
                          dtype=torch.float64, device=device)
  if world > 1:
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    dist.all_reduce(expand_s, op=dist.ReduceOp.MAX)  # the ranks expand their rules concurrently
  nnz_total, nodes_total, terms_total, launches_total, expanded_total = [float(x) for x in sizes.tolist()]
  expand_s, expand_kernels_s = [float(x) for x in expand_s.tolist()]
  strong = args.scaling == 'strong'
  # weak: every rank evaluates its own rules on its own copy of the table; strong: one problem, one table
  table_passes = 1 if strong else world
  job_bytes = 28.0 * nnz_total + table_passes * (24.0 * n + 8.0 * n * (1.0 + 2.0 / max(args.size_a - 1, 1)))
  value = job_bytes / (ms_step * 1e-3) / 1e9

  # parity of the timed workload, part 1: every flux term adds +w and -w (tm.scm:1288-1291), so the
  # dy/dt the timed loop left behind must sum to zero up to rounding
  result = out if sharded is None else (sharded.out if args.exchange == 'peer' else out_full)[:n]
  torch.cuda.synchronize()
  flux_sum, flux_abs = float(result.sum().item()), float(result.abs().sum().item())
  parity = dict(sum_dy_dt=flux_sum, sum_abs_dy_dt=flux_abs, sum_rel=abs(flux_sum) / max(flux_abs, 1e-300),
                sum_tolerance=1e-12, sum_ok=bool(abs(flux_sum) <= 1e-12 * flux_abs and flux_abs > 0))

  # per-phase device times of this rank's kernels (CUDA events on the launching stream)
  phase = numpy.zeros(3)
  reps = max(3, min(args.steps, 10))
  for _ in range(reps):
    phase += model.rhs_profile(p, out)
  phase /= reps
  peak, peak_src = measured_peak()
  sliced = info.get('flux_format', 0) == 1
  flux_kernel = 'flux_slices_kernel' if sliced else 'spmv_kernel'
  flux_bytes = flux_format_bytes(info, n) if sliced else spmv_bytes(info['nnz'], n)
  flux_gbs = flux_bytes / (phase[2] * 1e-3) / 1e9
  traffic = recorded_traffic(flux_kernel, info, args)
  roofline = dict(bound='hbm', kernel=flux_kernel + ' (S*w, one launch per step)', achieved=flux_gbs, peak=peak,
                  unit='GB/s', frac=flux_gbs / peak, traffic=traffic,
                  dram_frac=(traffic / (phase[2] * 1e-3) / 1e9 / peak) if traffic else None,
                  peak_source=peak_src,
                  algorithmic_bytes_per_launch=flux_bytes,
                  algorithmic_bytes_are=('compulsory bytes of the shipped format, each byte once: slice pointers + structure '
                                         'words + one 8-byte weight per stored term + per prefix group its list entry and sum '
                                         '+ ratio table + result (bench.py flux_format_bytes, DESIGN.md section 4); `traffic` '
                                         'above it = L2 misses of repeated reads' if sliced else
                                         'plain CSR: 12 * nnz + 16 * n'),
                  csr_equivalent_gbs=spmv_bytes(info['nnz'], n) / (phase[2] * 1e-3) / 1e9,
                  kernel_ms=float(phase[2]),
                  phases_ms=dict(marginals_and_world_probs=float(phase[0]), forest_levels=float(phase[1]),
                                 spmv=float(phase[2])))
  lv_gbs = level_bytes(info) / (phase[1] * 1e-3) / 1e9
  lv_parts = [recorded_traffic(k, info, args) for k in ('level_kernel', 'plane_kernel', 'prefix_sums_kernel')]
  lv_traffic = sum(x for x in lv_parts if x) if lv_parts[0] else None
  roofline_levels = dict(bound='hbm', kernel='plane_kernel + level_kernel per forest level and prefix_sums_kernel, summed',
                         achieved=lv_gbs, peak=peak, unit='GB/s', frac=lv_gbs / peak, traffic=lv_traffic,
                         dram_frac=(lv_traffic / (phase[1] * 1e-3) / 1e9 / peak) if lv_traffic else None,
                         algorithmic_bytes_per_step=level_bytes(info), kernel_ms=float(phase[1]))
  pre_bytes = prepass_bytes(info, n, args.size_a)
  step_traffic = recorded_traffic(None, info, args)  # every kernel of one right-hand side
  pre_traffic = (step_traffic - lv_traffic - traffic) if (step_traffic and lv_traffic and traffic) else None
  roofline_prepass = dict(bound='hbm', kernel='marginal tables + per-step ratio tables + leaf-world probabilities',
                          achieved=pre_bytes / (phase[0] * 1e-3) / 1e9, peak=peak, unit='GB/s',
                          frac=pre_bytes / (phase[0] * 1e-3) / 1e9 / peak, traffic=pre_traffic,
                          dram_frac=(pre_traffic / (phase[0] * 1e-3) / 1e9 / peak) if pre_traffic else None,
                          algorithmic_bytes_per_step=pre_bytes, kernel_ms=float(phase[0]))
  # the whole step as a bandwidth: DRAM bytes of all its kernels (ncu) over the timed step
  value_dram = (step_traffic / (ms_step * 1e-3) / 1e9) if (step_traffic and world == 1) else None

  # the expansion, once per structure: SURVEY.md section 8(d) counts about 85 algorithmic bytes per
  # expanded state (48 B record written, 32 B of hash-slot traffic, the record read again shared by
  # its children); time = this rank's device expansion without the driver's allocation time
  expand_rank_s = max((timing['device_expand_ms'] - timing['expand_alloc_ms']) * 1e-3, 1e-9)
  expand_gbs = 85.0 * info['n_nodes'] / expand_rank_s / 1e9
  roofline_expand = dict(bound='hbm', kernel='classify_kernel + emit_kernel / emit_right_only_kernel + emit_groups_kernel + scan + key sort '
                         '(frontier expansion, once per structure; hash inserts and compaction are latency- and '
                         'atomics-bound rather than streaming)', achieved=expand_gbs, peak=peak, unit='GB/s',
                         frac=expand_gbs / peak, traffic=None, algorithmic_bytes_per_state=85.0,
                         states=info['n_nodes'], seconds=expand_rank_s)

  # end to end through the reference-facing C ABI call with pinned HOST buffers
  e2e = None
  if rank == 0 and world == 1:
    h_in = torch.empty(n, dtype=torch.float64).pin_memory()
    h_out = torch.empty(n, dtype=torch.float64).pin_memory()
    h_in.copy_(p)
    f = mt.get_dy_dt(tag=tag, size_a=args.size_a, cl_k=args.cl_k)
    a_tag = numpy.frombuffer(tag.encode() + b'\x00', dtype=numpy.uint8)
    call = lambda: mt.u_lib.c_compute_dy_dt(a_tag.ctypes.data, args.cl_k, 0, h_in.data_ptr(), h_out.data_ptr())
    call()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
      call()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.e2e_steps
    e2e = dict(value=step_bytes(info['nnz'], n, args.size_a) / (e2e_ms * 1e-3) / 1e9, unit=UNIT,
               h2d_bytes_per_step=8 * n, d2h_bytes_per_step=8 * n, ms_per_step=e2e_ms,
               api='c_compute_dy_dt (host buffers, pinned)')
    # the same call the way a user of the reference module makes it: get_dy_dt(...)(numpy array, t)
    # with a PAGEABLE input array, a fresh result array per call (framework/markov_tapes.py:275-288)
    def drop_in():
      y = h_in.numpy().copy()  # pageable
      res = f(y, 0.0)
      res = f(y, 0.0)  # as in the loop below, the previous result is still referenced while the next is made
      t1 = time.perf_counter()
      for _ in range(args.e2e_steps):
        res = f(y, 0.0)
      ms = 1e3 * (time.perf_counter() - t1) / args.e2e_steps
      ok = bool(numpy.array_equal(res, h_out.numpy()))
      return dict(value=step_bytes(info['nnz'], n, args.size_a) / (ms * 1e-3) / 1e9, unit=UNIT, ms_per_step=ms,
                  api='markov_tapes.get_dy_dt(...)(pageable numpy array, t) -> new numpy array',
                  same_bits_as_pinned_call=ok)
    e2e['drop_in'] = _guarded(drop_in)
    del f
  elif world > 1:
    # N > 1: the table comes in once over the N PCIe links together - rank r copies its 1 / N slice from
    # pinned host memory and the ranks all-gather the slices over NVLink - the ranks evaluate the step
    # together, and rank 0 takes the summed dy/dt back out to pinned host memory (every rank holds it).
    # Nothing inside the guarded part but the step's own collectives (whose waits time out or belong to
    # NCCL's watchdog), and the figures meet in one all-reduce afterwards.
    mine_ms, failed = 0.0, 0.0
    share = -(-n // world)
    try:
      h_in = torch.empty(share * world, dtype=torch.float64).pin_memory()
      h_in[:n].copy_(p)
      h_out = torch.empty(n, dtype=torch.float64).pin_memory() if rank == 0 else None
      table = torch.zeros(max(share * world, p_full.numel()), dtype=torch.float64, device=device)
      mine = torch.empty(share, dtype=torch.float64, device=device)

      def e2e_step():
        mine.copy_(h_in[rank * share:(rank + 1) * share], non_blocking=True)
        dist.all_gather_into_tensor(table[:share * world], mine)
        one_step(table)
        if rank == 0:
          result = sharded.out if args.exchange == 'peer' else out_full
          h_out.copy_(result[:n], non_blocking=True)
        torch.cuda.synchronize()

      e2e_step()
      dist.barrier()
      t0 = time.perf_counter()
      for _ in range(args.e2e_steps):
        e2e_step()
      mine_ms = 1e3 * (time.perf_counter() - t0) / args.e2e_steps
    except Exception as ex:  # pylint: disable=broad-except
      failed = 1.0
      print(f'rank {rank}: end-to-end measurement failed: {ex!r}', file=sys.stderr)
    both = torch.tensor([mine_ms, failed], dtype=torch.float64, device=device)
    dist.all_reduce(both, op=dist.ReduceOp.MAX)
    e2e_ms, any_failed = float(both[0].item()), float(both[1].item()) > 0
    if rank == 0:
      if any_failed or not e2e_ms > 0:
        e2e = dict(value=None, unit=UNIT, h2d_bytes_per_step=8 * n, d2h_bytes_per_step=8 * n,
                   note='end-to-end measurement failed on some rank (see stderr)')
      else:
        e2e = dict(value=job_bytes / (e2e_ms * 1e-3) / 1e9, unit=UNIT, h2d_bytes_per_step=8 * share * world,
                   d2h_bytes_per_step=8 * n, ms_per_step=e2e_ms,
                   api='pinned host table -> 1 / N slice per rank (H2D) + all-gather over NVLink, step of all ranks '
                       'together, summed dy/dt -> pinned host on rank 0 (D2H); max over ranks')

  # N > 1, weak run: the strong-scaling figure beside it.  The N = 1 problem (rank 0's 24 rules, which
  # is also what every rank evaluates alone in the weak run, so its one-GPU time is rank_compute_ms) is
  # dealt to the N ranks by term counts and evaluated with the same fused exchange.
  strong_leg = None
  if world > 1 and args.scaling == 'weak' and args.exchange == 'peer' and not args.no_strong_leg:
    def measure_strong():
      base = configs.random_rule_set(args.size_a, args.rules_per_gpu, seed=args.seed)
      part_tag = configs.synthetic_tag(args.size_a, args.rules_per_gpu, args.seed) + f'-part{rank}of{world}'
      built = 1.0
      try:
        mt.register_rule_set(part_tag, args.size_a, parallel.split_rule_set(base, world, rank, args.size_a, args.cl_k))
        part_model = dev.DeviceModel(part_tag, args.cl_k)
      except Exception as ex:  # pylint: disable=broad-except
        built = 0.0
        print(f'rank {rank}: strong-scaling structure failed: {ex!r}', file=sys.stderr)
      flag = torch.tensor([built], dtype=torch.float64, device=device)
      dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # every rank reaches this, so nobody waits for a rank that failed
      if float(flag.item()) == 0:
        return dict(error='building a share of the rules failed on some rank (see stderr)')
      ex_strong = parallel.PeerExchangeRhs(part_model, rounds=max(args.chunks, 1) if args.chunks <= 16 else 16)
      for _ in range(3):
        ex_strong.rhs_full(p_full)
      torch.cuda.synchronize()
      dist.barrier()
      s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s0.record()
      for _ in range(10):
        ex_strong.rhs_full(p_full)
      s1.record()
      torch.cuda.synchronize()
      t = torch.tensor([s0.elapsed_time(s1) / 10], dtype=torch.float64, device=device)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      bad = torch.tensor([float(ex_strong.lib.tapes_peer_group_error(ex_strong.group))], dtype=torch.float64, device=device)
      dist.all_reduce(bad, op=dist.ReduceOp.MAX)  # all ranks take the same way out
      got = ex_strong.out[:n]
      total, gross = float(got.sum().item()), float(got.abs().sum().item())
      ex_strong.close()
      mt.u_lib.tapes_release_model(part_tag.encode(), args.cl_k)
      if float(bad.item()) != 0:
        return dict(error='a cross-GPU wait timed out')
      one_gpu = max(rank_ms)
      return dict(ms_per_step=float(t.item()), one_gpu_ms=one_gpu, speedup=one_gpu / float(t.item()), n_gpus=world,
                  rules=args.rules_per_gpu, sum_dy_dt=total, sum_abs_dy_dt=gross, sum_rel=abs(total) / max(gross, 1e-300),
                  note='the N = 1 problem dealt to the N ranks (bench.py --scaling strong times the same thing as its own '
                       'line); one_gpu_ms = this run\'s time of the same rules on one GPU')
    from chemical_kinetics_and_program_execution_b200 import configs
    strong_leg = _guarded(measure_strong)

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    from oracle import oracle
    oracle.build()
    # the sample is the whole timed workload unless --cpu-rules asks for fewer rules: all rules on
    # the full table, so that its dy/dt can be held against the GPU's (parity of the timed workload,
    # part 2); fewer rules are compared with a GPU structure built for exactly those rules
    n_sample = min(args.cpu_rules or args.rules_per_gpu, args.rules_per_gpu)
    sample = {k: numpy.asarray(v)[:n_sample] for k, v in rules.items()}
    t_cpu, cores, counters, cpu_out = cpu_port_step(args, sample, want_result=True)
    nnz_s = 2 * counters['acc_calls']
    cpu = dict(value=step_bytes(nnz_s, n, args.size_a) / t_cpu / 1e9, unit=UNIT, cores=cores, kind='port',
               sample=(f'{"all" if n_sample == args.rules_per_gpu else "first"} {n_sample} of {args.rules_per_gpu} rules '
                       f'on the full {n}-state table '
                       f'(nnz={nnz_s}), merged-mode CPU port of compute-dy/dt, {cores} worker processes, '
                       f'{t_cpu:.1f} s'),
               seconds=t_cpu, states_expanded_per_s=(counters['ext_nodes'] + counters['worlds']) / t_cpu,
               reference_equivalent_work=_guarded(literal_vs_merged, args, rules))

    def against_cpu():
      if n_sample == args.rules_per_gpu:
        gpu_out = out.cpu().numpy()
      else:
        sample_tag = tag + f'-first{n_sample}'
        mt.register_rule_set(sample_tag, args.size_a, sample)
        sample_model = dev.DeviceModel(sample_tag, args.cl_k)
        gpu_out = sample_model.rhs(p).cpu().numpy()
        mt.u_lib.tapes_release_model(sample_tag.encode(), args.cl_k)
      scale = float(abs(cpu_out).max())
      err = float(abs(gpu_out - cpu_out).max())
      return dict(rules_compared=n_sample, states_compared=int(n), max_abs_err=err, max_abs_dy_dt=scale,
                  max_rel_err=err / max(scale, 1e-300), tolerance=1e-12, ok=bool(err <= 1e-12 * scale and scale > 0),
                  note='GPU dy/dt of the timed workload against the CPU port on the same rules and table, '
                       'error relative to the largest |dy/dt|')
    parity['vs_cpu_port'] = _guarded(against_cpu)
    del cpu_out

  if rank == 0:
    parity['ok'] = bool(parity['sum_ok'] and parity.get('vs_cpu_port', {}).get('ok', True))
    exchange = None
    if world > 1:
      exchange = ('fused into the product kernel: partial flux stored into the owner\'s slots over NVLink peer memory, '
                  f'owners sum and broadcast one round behind ({max(args.chunks, 1)} rounds)' if args.exchange == 'peer' else
                  f'all-reduce of dy/dt in {max(args.chunks, 1)} row blocks overlapped with the product'
                  if args.exchange == 'allreduce' else
                  f'flux reduce-scatter + table all-gather, {args.chunks} overlapped row chunks')
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_step, higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype='f64',
                data='synthetic', config=workload_config(args, world),
                value_is='CSR-equivalent algorithmic bytes per second (a work rate; may exceed the HBM peak)',
                value_dram_gbs=value_dram, dram_frac=(value_dram / peak) if value_dram else None,
                structure=dict(nnz=nnz_total, forest_nodes=nodes_total, flux_terms=terms_total, exchange=exchange),
                parity_check=parity, strong_scaling=strong_leg,
                clocks=clocks.summary(), e2e=e2e,
                gpu_launches=(int(launches_total / world)
                              + (4 * max(args.chunks, 1) + 1 if world > 1 and args.exchange == 'peer' else 0)) * args.steps,
                roofline=roofline, roofline_levels=roofline_levels, roofline_prepass=roofline_prepass,
                roofline_expand=roofline_expand, cpu_baseline=cpu, rank_compute_ms=rank_ms, exchange_ms=comm_ms,
                exchange_exposed_ms=(ms_step - max(rank_ms)) if rank_ms else None,
                states_expanded_per_s=expanded_total / max(expand_kernels_s, 1e-9),
                states_expanded_per_s_incl_driver_alloc=expanded_total / max(expand_s, 1e-9),
                states_expanded_note=('forest nodes + program worlds of all ranks / (host enumeration + device expansion) of '
                                      'the slowest rank; the first figure leaves out the time inside cudaMalloc / cudaFree '
                                      '(build.expand_alloc_ms), the second includes it; before the build is timed the '
                                      'device memory is touched once (build.prewarm; --no-prewarm to leave that out), '
                                      'neither figure contains that'),
                build=dict(seconds=build_s, **timing, prewarm=prewarm, forest_levels=info['n_levels'],
                           hash_inserts=info['hash_inserts'], hash_unique=info['hash_unique'],
                           flux_slices={k: info.get(k) for k in ('n_slices', 'slice_words', 'runs', 'run_entries',
                                                                  'column_entries', 'column_slots')}))
    emit(line)
  if world > 1:
    dist.destroy_process_group()


def main():
  args = parse_args()
  claim_stdout()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
