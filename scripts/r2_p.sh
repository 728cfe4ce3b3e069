set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_p_device.txt
# Msp solve: every solver configuration at 1024^2 (one matrix), tuned plan printed for the default
timeout 90 python scripts/probe_msp.py 1024 --configs 1:0:0,2:0:0,2:1:0,2:1:1 > gpurun_out/r2_p_msp_1024.log 2>&1; tail -4 gpurun_out/r2_p_msp_1024.log
# the whole GPU suite (4 workers: most of the wall time is the CPU oracle)
timeout 400 python -m pytest tests -m gpu -q -n 4 --durations=12 > gpurun_out/r2_p_gpu_tests.log 2>&1; tail -25 gpurun_out/r2_p_gpu_tests.log
# the driver's bench line
timeout 240 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_p_bench_n1.json 2> gpurun_out/r2_p_bench_n1.err; echo bench rc=$?; tail -c 600 gpurun_out/r2_p_bench_n1.err
# ncu launch list of the headline command (kernel share of the step)
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_p_launches.csv python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_p_ncu.log 2>&1; tail -2 gpurun_out/r2_p_ncu.log
