timeout 150 python -u -m pytest tests -m gpu -v -x -p no:cacheprovider --deselect tests/test_gpu_msp.py 2>&1 | while IFS= read -r l; do printf '%s %s\n' "$(date +%s.%N | cut -c1-14)" "$l"; done > gpurun_out/r2_r_serial_tests.log
tail -5 gpurun_out/r2_r_serial_tests.log
