TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR scripts/check_multi_gpu.py > gpurun_out/r02_g_check_multi_gpu_n2.log 2>&1; echo "check rc=$?"; tail -12 gpurun_out/r02_g_check_multi_gpu_n2.log
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_g_bench_n2_weak.json 2> gpurun_out/r02_g_bench_n2_weak.err; echo "weak rc=$?"; tail -2 gpurun_out/r02_g_bench_n2_weak.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong > gpurun_out/r02_g_bench_n2_strong.json 2> gpurun_out/r02_g_bench_n2_strong.err; echo "strong rc=$?"; tail -2 gpurun_out/r02_g_bench_n2_strong.err
python - <<'PY'
import json
for f in ('weak','strong'):
    try:
        d=json.load(open(f'gpurun_out/r02_g_bench_n2_{f}.json'))
        print(f, 'ms', d['ms_per_step'], 'value', d['value'], 'rank_ms', d['rank_compute_ms'], 'exposed', d['exchange_exposed_ms'], 'e2e', d['e2e'], 'parity', d['parity_check'])
    except Exception as ex: print(f, ex)
PY
