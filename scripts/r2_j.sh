set -x
timeout 900 python -m pytest tests/test_gpu_apply3d.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2_j_gpu.log 2>&1; tail -8 gpurun_out/r2_j_gpu.log
python scripts/probe3d.py 256 512 > gpurun_out/r2_j_probe3d.log 2>&1; cat gpurun_out/r2_j_probe3d.log
