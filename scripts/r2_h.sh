set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi -L > gpurun_out/r2_h_gpus.txt
timeout 900 python -m pytest tests/test_dist.py -x -q -m gpu -k "nccl and 8" > gpurun_out/r2_h_dist8.log 2>&1; tail -5 gpurun_out/r2_h_dist8.log
timeout 300 $TR --master-port 29544 scripts/probe3d_dist.py 256 1 2 4 2>&1 | grep "^P=" | tee gpurun_out/r2_h_probe256.log
LS_PROBE_BEST=gpurun_out/r2_h_best512.txt timeout 400 $TR --master-port 29541 scripts/probe3d_dist.py 512 1 2 4 8 2>&1 | grep "^P=" | tee gpurun_out/r2_h_probe512.log
LS_OP3D_XCHG=nccl timeout 300 $TR --master-port 29542 scripts/probe3d_dist.py 512 1 4 2>&1 | grep "^P=" | sed 's/^/nccl /' | tee -a gpurun_out/r2_h_probe512.log
LS_OP3D_XCHG=nccl timeout 300 $TR --master-port 29543 scripts/probe3d_dist.py 256 1 2>&1 | grep "^P=" | sed 's/^/nccl /' | tee -a gpurun_out/r2_h_probe256.log
timeout 900 $TR --master-port 29545 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_h_bench_n8.json 2> gpurun_out/r2_h_bench_n8.err; echo rc=$?; tail -c 1500 gpurun_out/r2_h_bench_n8.err
