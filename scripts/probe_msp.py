"""Device Msp^-1 at large grids: factorisation time / memory, solve time, residual (random 9-point matrix)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fast_solver_lippmann_schwinger_b200 as ls
from test_gpu_msp import stencil9

for n in [int(a) for a in sys.argv[1:]] or [512, 1024]:
    t0 = time.perf_counter()
    A = stencil9(n, n, seed=1)
    t1 = time.perf_counter()
    F = ls.GPUMspFactorization(A, n, n)
    t2 = time.perf_counter()
    N = n * n
    rng = np.random.default_rng(0)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
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
    print("n=%d build %.1fs factor(total call) %.2fs (inside %.2fs) depth %d factor %.2f GB  solve %.3f ms -> %.0f GB/s  residual %.2e" % (
        n, t1 - t0, t2 - t1, F.factor_seconds, F.depth, F.factor_bytes / 1e9, ms, F.factor_bytes / ms / 1e6, res), flush=True)
    F.destroy()
