#!/bin/bash
# Round-2 evidence on one B200 (files land in gpurun_out/, copied to profiles/ afterwards).
python -m pytest tests -m gpu -q > gpurun_out/r02_z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_z_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_z_bench_n1e8.json 2> gpurun_out/r02_z_bench_n1e8.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r02_z_bench_reference_n1e8_2steps.json 2> gpurun_out/r02_z_bench_reference.err; echo "ref rc=$?"
python scripts/time_small_r02.py > gpurun_out/r02_z_time_small.log 2>&1; echo "small rc=$?"
(cd examples && for f in ex2_ferromagnet ex4_chemical_turing ex5_msrtf_machine own_program; do echo "== $f"; python $f.py; done; echo "== ex2_ferromagnet_mc (100 trials)"; rm -f ferromagnet_mc_chain_counts.npz; python ex2_ferromagnet_mc.py) > gpurun_out/r02_z_examples.log 2>&1; echo "examples rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_z_launches_bench_n1e8_24rules.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
python scripts/one_rhs.py 10 8 24 2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"level_kernel|plane_kernel|flux_slices|ratio|marginal|rule_weight|root_kernel|prefix_sums" -s 29 -c 29 -o gpurun_out/r02_z_step python scripts/one_rhs.py 10 8 24 2 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log
