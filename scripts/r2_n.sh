set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_n_gpu.log 2>&1; tail -5 gpurun_out/r2_n_gpu.log
python scripts/probe3d_rect.py 512 512 > gpurun_out/r2_n_rect.log 2>&1; cat gpurun_out/r2_n_rect.log
python scripts/probe2d.py 512 > gpurun_out/r2_n_2d.log 2>&1; cat gpurun_out/r2_n_2d.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_n_bench_n1.json 2> gpurun_out/r2_n_bench_n1.err; echo rc=$?; tail -c 800 gpurun_out/r2_n_bench_n1.err
