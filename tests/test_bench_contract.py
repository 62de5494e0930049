"""CPU checks of bench.py's contract: the reference arm prints exactly one JSON line with the keys
the driver reads, and the helper figures are what DESIGN.md says they are."""

import json
import os
import subprocess
import sys
import types

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2',
                        '--warmup', '1', '--size-a', '3', '--cl-k', '5', '--rules-per-gpu', '3'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert out.returncode == 0, out.stderr
  lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
  assert len(lines) == 1
  line = json.loads(lines[0])
  assert line['impl'] == 'reference' and line['steps'] == 2 and line['warmup'] == 1 and line['n_gpus'] == 1
  assert line['higher_is_better'] is True and line['unit'] == 'GB/s' and line['value'] > 0
  assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
  assert line['cpu_baseline']['value'] == line['value']
  assert line['e2e'] == dict(value=line['value'], unit='GB/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
  assert line['config']['workload'] == 'synthetic-random-rewrite-rules'
  # the reference arm runs the bench's own configuration (it used to shorten cl_k and
  # still report the bench's): `config` is the object this repo's arm prints for the same flags
  sys.path.insert(0, ROOT)
  import bench
  args = types.SimpleNamespace(size_a=3, cl_k=5, rules_per_gpu=3, seed=1, scaling='weak')
  assert line['config'] == bench.workload_config(args, 1)
  assert line['config']['cl_k'] == 5 and line['config']['n_states'] == 3 ** 5 and line['same_config'] is True
  assert 'cl_k=5' in line['cpu_baseline']['sample'] and '243-state' in line['cpu_baseline']['sample']


def test_reference_arm_says_what_it_ran_when_the_budget_is_short():
  """With a budget that cannot hold warmup + steps evaluations the warm-up goes first, then steps;
  the line reports what was run, on the same configuration (never a shorter window)."""
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '4',
                        '--warmup', '3', '--size-a', '3', '--cl-k', '6', '--rules-per-gpu', '3', '--ref-budget-s', '0'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert out.returncode == 0, out.stderr
  line = json.loads(out.stdout.strip())
  assert line['steps'] == 1 and line['warmup'] == 0 and line['steps_requested'] == 4 and line['warmup_requested'] == 3
  assert line['config']['cl_k'] == 6 and line['config']['n_states'] == 3 ** 6


def test_cpu_port_step_returns_the_dy_dt_of_the_whole_rule_set():
  """The in-bench parity check compares the GPU's dy/dt with what the CPU workers add up to."""
  sys.path.insert(0, ROOT)
  import bench
  from chemical_kinetics_and_program_execution_b200 import configs
  from oracle import oracle
  oracle.build()
  args = types.SimpleNamespace(size_a=3, cl_k=4, seed=5)
  rules = configs.random_rule_set(3, 5, seed=9)
  seconds, workers, counters, total = bench.cpu_port_step(args, rules, want_result=True)
  oracle.register_rules('whole-set', 3, rules)
  p = configs.product_table(numpy.random.default_rng(7).dirichlet(numpy.ones(3)), 4)
  want = oracle.compute_dy_dt('whole-set', 4, p, mode=oracle.MERGED)
  assert workers >= 1 and seconds > 0 and counters['acc_calls'] > 0
  assert abs(total - want).max() <= 1e-15 * abs(want).max()
  assert abs(total.sum()) <= 1e-12 * abs(total).sum()


def test_other_ranks_of_the_reference_arm_stay_silent():
  env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                        '--steps', '1', '--warmup', '0', '--size-a', '3', '--cl-k', '4', '--rules-per-gpu', '2'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
  assert out.returncode == 0 and out.stdout.strip() == ''


def test_byte_counts_and_reference_equivalent_work():
  sys.path.insert(0, ROOT)
  import bench
  from chemical_kinetics_and_program_execution_b200 import configs
  assert bench.spmv_bytes(10, 3) == 12 * 10 + 16 * 3
  assert bench.step_bytes(10, 3, 5) == 28 * 10 + 24 * 3 + 8 * 3 * 1.5
  info = dict(n_slices=2, slice_words=100, n_terms=30, nnz=60)
  assert bench.flux_format_bytes(info, 40) == 8 * 3 + 4 * 2 + 4 * 100 + 8 * 30 + 8 * 40
  grouped = dict(info, nnz_stored=20, hash_unique=2, alphabet=10)
  # grouped right children: list entry + sum per group, the ratio table once, offset + outflow sum per prefix
  assert bench.flux_format_bytes(grouped, 40) == 8 * 3 + 4 * 2 + 4 * 100 + 8 * 10 + 16 * 2 + 8 * 40 + 16 * 4 + 8 * 40
  levels = dict(n_nodes=1000, alphabet=10, hash_inserts=300, owned_parents=100, plane_groups=0, hash_unique=50, n_states=100,
                materialize_right=1, left_parents=20, deferred_groups=5, distinct_table_reads=400)
  assert bench.level_bytes(levels) == 8 * 400 + 8 * 1000 + 25 * 20 + 16 * 50 + 8 * 200 + (8 * 50 + 8 * 5 + 12 * 50 + 16 * 10)
  assert bench.level_bytes(dict(levels, distinct_table_reads=1000)) - bench.level_bytes(levels) == 8 * 600
  work = bench.literal_vs_merged(types.SimpleNamespace(size_a=4, cl_k=7), configs.random_rule_set(4, 6, seed=2))
  rows = work['counted']
  assert rows[0]['cl_k'] == 2 and all(r['literal_nodes'] >= r['merged_nodes'] for r in rows)
  assert rows[-1]['literal_accumulate_calls'] > rows[-1]['merged_terms']  # windows revisited per left context
  if 'extrapolated' in work:
    assert work['extrapolated']['literal_nodes'] > work['extrapolated']['merged_nodes']
  assert 'error' in bench._guarded(bench.literal_vs_merged, types.SimpleNamespace(size_a=4, cl_k=7), {'bad': 1})
