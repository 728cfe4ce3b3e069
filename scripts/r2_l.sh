set -x
V=fast_solver_lippmann_schwinger_b200/lib/variants
for v in "" $V/libls_cuda_lb4_inv2.so $V/libls_cuda_lb8_inv3.so $V/libls_cuda_lb8_inv2.so; do echo "== lib ${v:-default(lb4_inv3)}"; LS_CUDA_LIB=$v python scripts/probe3d.py 512; done > gpurun_out/r2_l_512_variants.log 2>&1; cat gpurun_out/r2_l_512_variants.log
for v in "" $V/libls_cuda_lb4_inv2.so; do echo "== lib ${v:-default}"; LS_CUDA_LIB=$v python scripts/probe2d.py 512 1024; done > gpurun_out/r2_l_2d.log 2>&1; cat gpurun_out/r2_l_2d.log
for cw in 2 8; do LS_C_INTERLEAVE=$cw python scripts/probe2d_cw.py 2048 $cw; done > gpurun_out/r2_l_cw.log 2>&1; cat gpurun_out/r2_l_cw.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_msp_g -s 400 -c 76 --csv --log-file gpurun_out/r2_l_msp_launches.csv python scripts/probe_msp.py 2048 > gpurun_out/r2_l_msp_ncu.log 2>&1; tail -2 gpurun_out/r2_l_msp_ncu.log
