set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu -k "nccl and 8" > gpurun_out/r2_o_dist8.log 2>&1; tail -4 gpurun_out/r2_o_dist8.log
timeout 200 $TR --master-port 29544 scripts/probe3d_dist.py 256 1 2>&1 | grep "^P=" | sed 's/^/auto /' | tee gpurun_out/r2_o_probe256.log
LS_OP3D_XCHG=ce timeout 200 $TR --master-port 29543 scripts/probe3d_dist.py 256 1 2 2>&1 | grep "^P=" | sed 's/^/ce+flags /' | tee -a gpurun_out/r2_o_probe256.log
timeout 300 $TR --master-port 29541 scripts/probe3d_dist.py 512 2 4 2>&1 | grep "^P=" | sed 's/^/auto(ce+flags) /' | tee gpurun_out/r2_o_probe512.log
