set -x
timeout 900 python -m pytest tests/test_gpu_apply2d.py tests/test_gpu_apply3d.py tests/test_gpu_msp.py -x -q > gpurun_out/r2_g_gpu.log 2>&1; tail -8 gpurun_out/r2_g_gpu.log
for c in 2 4 8 16; do LS_MSP_COLS_PER_LANE=$c timeout 300 python scripts/probe_msp.py 2048; done > gpurun_out/r2_g_msp.log 2>&1; cat gpurun_out/r2_g_msp.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_g_bench_n1.json 2> gpurun_out/r2_g_bench_n1.err; echo rc=$?; tail -c 1500 gpurun_out/r2_g_bench_n1.err
