"""Timing probe of the sharded 3-D apply under torchrun: serial exchange (1 chunk) vs pipelined x-slot chunks.
   torchrun --nproc-per-node P scripts/probe3d_dist.py N [chunks ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import fast_solver_lippmann_schwinger_b200 as ls
from fast_solver_lippmann_schwinger_b200 import dist as lsd
from fast_solver_lippmann_schwinger_b200._lib import check, lib
from fast_solver_lippmann_schwinger_b200.problems import nu_gaussian_3d_grid

rank, world, local = lsd.env_rank()
torch.cuda.set_device(local)
check(lib().ls_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunks = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
h = 1.0 / n; k = 2 * np.pi / (10 * h)
a, b_ = lsd.vector_range(n, n, n, rank, world)
nu = nu_gaussian_3d_grid(n)[a:b_]
rng = np.random.default_rng(4321 + rank)
b = rng.standard_normal(b_ - a) + 1j * rng.standard_normal(b_ - a)
db = ls.DeviceBuffer.from_host(b); dy = ls.DeviceBuffer(b.nbytes)
ref = None
best = (1e9, 1)
for cx in chunks:
    os.environ["LS_OP3D_CHUNKS"] = str(cx)
    uid = lsd.broadcast_unique_id(rank)
    M = lsd.FastM3DSharded(nu, n, n, n, k, 1.8 * n * h, 4.0 * n * h, rank, world, uid)
    for _ in range(3):
        M.mul_(dy, db)
    M.sync(); dist.barrier()
    y = dy.to_host(np.complex128, b_ - a) if hasattr(dy, "to_host") else None
    reps = 10
    M.profile_enable(True)
    M.sync(); dist.barrier()
    M.timer_start()
    for _ in range(reps):
        M.mul_(dy, db)
    ms = M.timer_stop() / reps
    ph, cnt = M.profile_read(7)
    M.profile_enable(False)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    same = None
    if y is not None:
        same = True if ref is None else bool(np.array_equal(y, ref))
        if ref is None:
            ref = y
    if rank == 0:
        per = [p / reps for p in ph]
        print("P=%d n=%d chunks=%d apply %.3f ms (max over ranks) -> %.1f applies/s; compute %.3f a2a %.3f+%.3f phases %s same_bits=%s" % (
            world, n, cnt[1] // reps, float(t[0]), 1e3 / float(t[0]), sum(per[:5]), per[5], per[6], ["%.3f" % p for p in per], same), flush=True)
    best = min(best, (float(t[0]), cx))
    M.destroy()
if rank == 0 and os.environ.get("LS_PROBE_BEST"):
    open(os.environ["LS_PROBE_BEST"], "w").write(str(best[1]))
dist.barrier()
dist.destroy_process_group()
