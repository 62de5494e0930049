python -m pytest tests -m gpu -q > gpurun_out/r02_z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_z_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_z_bench_n1e8.json 2> gpurun_out/r02_z_bench_n1e8.err; echo "bench rc=$?"
python scripts/time_small_r02.py > gpurun_out/r02_z_time_small.log 2>&1; echo "small rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_z_launches_bench_n1e8_24rules.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
