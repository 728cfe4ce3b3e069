set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_k_gpu.log 2>&1; tail -8 gpurun_out/r2_k_gpu.log
python scripts/probe3d.py 512 > gpurun_out/r2_k_probe3d.log 2>&1; cat gpurun_out/r2_k_probe3d.log
python scripts/probe2d.py 512 > gpurun_out/r2_k_probe2d.log 2>&1; cat gpurun_out/r2_k_probe2d.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_msp -c 160 --csv --log-file gpurun_out/r2_k_msp_launches.csv python scripts/probe_msp.py 2048 > gpurun_out/r2_k_msp_ncu.log 2>&1; tail -2 gpurun_out/r2_k_msp_ncu.log
