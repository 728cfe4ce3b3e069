set -x
timeout 160 python -m pytest tests/test_gpu_msp.py -q --durations=10 -k "variants or tight or leaf or superlu or rejects" > gpurun_out/r2_q_msp_tests.log 2>&1; tail -22 gpurun_out/r2_q_msp_tests.log
