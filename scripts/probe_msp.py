"""Device Msp^-1 at large grids: factorisation time / memory, solve time, residual (random 9-point matrix), for every
solver configuration:  python scripts/probe_msp.py 2048 [--configs 1:0:0,2:0:0,2:1:0,2:1:1] [--plan]
(config = LS_MSP_SOLVER:LS_MSP_FUSE:LS_MSP_TUNE; the matrix is built once per grid)."""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fast_solver_lippmann_schwinger_b200 as ls
from test_gpu_msp import stencil9

args = sys.argv[1:]
configs = None
show_plan = "--plan" in args
if "--configs" in args:
    configs = args[args.index("--configs") + 1].split(",")
sizes = [int(a) for a in args if a.isdigit()] or [512, 1024]
if configs is None:
    configs = [os.environ.get("LS_MSP_SOLVER", "2") + ":" + os.environ.get("LS_MSP_FUSE", "1") + ":" + os.environ.get("LS_MSP_TUNE", "1")]

for n in sizes:
    t0 = time.perf_counter()
    A = stencil9(n, n, seed=1)
    t1 = time.perf_counter()
    N = n * n
    rng = np.random.default_rng(0)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    for cfg in configs:
        so, fu, tu = cfg.split(":")
        os.environ["LS_MSP_SOLVER"], os.environ["LS_MSP_FUSE"], os.environ["LS_MSP_TUNE"] = so, fu, tu
        t2 = time.perf_counter()
        F = ls.GPUMspFactorization(A, n, n)
        t3 = time.perf_counter()
        db = ls.DeviceBuffer.from_host(b); dx = ls.DeviceBuffer(b.nbytes)
        for _ in range(3):
            F.solve(db, dx)
        F.sync()
        F.timer_start()
        reps = 10
        for _ in range(reps):
            F.solve(db, dx)
        ms = F.timer_stop() / reps
        x = dx.to_host()
        res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
        plan = F.plan()
        print("n=%d cfg solver:fuse:tune=%s build %.1fs factor(total call) %.2fs (inside %.2fs) depth %d factor %.3f GB  solve %.3f ms -> %.0f GB/s  "
              "residual %.2e  sha1 %s  [%s]" % (n, cfg, t1 - t0, t3 - t2, F.factor_seconds, F.depth, F.factor_bytes / 1e9, ms,
                                                F.factor_bytes / ms / 1e6, res, hashlib.sha1(x.tobytes()).hexdigest()[:12],
                                                plan.splitlines()[0]), flush=True)
        if show_plan:
            print(plan, flush=True)
        db.free(); dx.free()
        F.destroy()
