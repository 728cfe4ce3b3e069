TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_dist.py -x -q -m gpu -k "nccl and 8" 2>&1 | tail -5
LS_PROBE_BEST=gpurun_out/best512.txt timeout 300 $TR --master-port 29541 scripts/probe3d_dist.py 512 1 2 4 8 2>&1 | grep "^P=" 
NCCL_MAX_NCHANNELS=8 timeout 200 $TR --master-port 29542 scripts/probe3d_dist.py 512 4 2>&1 | grep "^P=" | sed 's/^/maxch8 /'
LS_OP3D_COMM_PRIO=0 timeout 200 $TR --master-port 29543 scripts/probe3d_dist.py 512 4 2>&1 | grep "^P=" | sed 's/^/prio0 /'
timeout 200 $TR --master-port 29544 scripts/probe3d_dist.py 256 1 2 4 2>&1 | grep "^P="
LS_OP3D_CHUNKS=$(cat gpurun_out/best512.txt) timeout 400 $TR --master-port 29545 bench.py --gpus 8 --steps 20 --warmup 3 --n3 512 > gpurun_out/bench_r1_i_n8_512.json 2> gpurun_out/bench_r1_i_n8_512.err; tail -c 1500 gpurun_out/bench_r1_i_n8_512.json
