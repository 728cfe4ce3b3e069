set -x
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q -k "nccl and 2" > gpurun_out/r2_f_dist2.log 2>&1; tail -15 gpurun_out/r2_f_dist2.log
timeout 600 python scripts/probe_msp.py 1024 2048 > gpurun_out/r2_f_msp.log 2>&1; cat gpurun_out/r2_f_msp.log
timeout 300 python -m pytest tests/test_gpu_msp.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_f_bench_n2.json 2> gpurun_out/r2_f_bench_n2.err; echo rc=$?; tail -c 2000 gpurun_out/r2_f_bench_n2.err
LS_OP3D_XCHG=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_f_bench_n2_nccl.json 2> gpurun_out/r2_f_bench_n2_nccl.err; echo rc=$?
