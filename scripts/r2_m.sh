set -x
python scripts/probe3d_rect.py 512 256 256 512 512 512 256 256 > gpurun_out/r2_m_rect.log 2>&1; cat gpurun_out/r2_m_rect.log
python scripts/probe_mgs.py 2048 > gpurun_out/r2_m_mgs.log 2>&1; cat gpurun_out/r2_m_mgs.log
for g in 0 1; do LS_MSP_GRAPH=$g python scripts/probe_msp.py 2048; done > gpurun_out/r2_m_msp.log 2>&1; cat gpurun_out/r2_m_msp.log
timeout 300 python -m pytest tests/test_gpu_msp.py -x -q 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_mid_fused -s 2 -c 1 -o gpurun_out/r2_m_p3_512 -f python scripts/probe3d_rect.py 512 512 > gpurun_out/r2_m_ncu.log 2>&1; tail -2 gpurun_out/r2_m_ncu.log
