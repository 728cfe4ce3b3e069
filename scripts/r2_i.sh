set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q -k "nccl and 2" > gpurun_out/r2_i_dist2.log 2>&1; tail -15 gpurun_out/r2_i_dist2.log
timeout 300 $TR --master-port 29544 scripts/probe3d_dist.py 256 1 4 2>&1 | grep "^P=" | tee gpurun_out/r2_i_probe.log
LS_OP3D_SYNC=barrier timeout 300 $TR --master-port 29546 scripts/probe3d_dist.py 256 1 4 2>&1 | grep "^P=" | sed 's/^/barrier /' | tee -a gpurun_out/r2_i_probe.log
timeout 600 python -m pytest tests/test_gpu_msp.py tests/test_gpu_zz_sparsifier3d.py tests/test_gpu_apply2d.py -x -q > gpurun_out/r2_i_gpu.log 2>&1; tail -8 gpurun_out/r2_i_gpu.log
