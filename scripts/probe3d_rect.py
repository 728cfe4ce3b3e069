"""3-D apply on an n x n x l grid (device-resident): per-phase times.  python scripts/probe3d_rect.py n l [n l ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fast_solver_lippmann_schwinger_b200 as ls

args = [int(a) for a in sys.argv[1:]]
for n, l in zip(args[0::2], args[1::2]):
    h = 1.0 / n
    k = 2 * np.pi / (10 * h)
    N = n * n * l
    nu = np.full(N, 0.1)
    M = ls.FastM3D(None, nu, 4 * n, 4 * n, 4 * l, n, n, l, k, L=1.8 * n * h, Lp=4.0 * n * h)
    rng = np.random.default_rng(1)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    db = ls.DeviceBuffer.from_host(b); dy = ls.DeviceBuffer(b.nbytes)
    for _ in range(2):
        M.mul_(dy, db)
    M.sync()
    reps = 5
    M.profile_enable(True)
    M.timer_start()
    for _ in range(reps):
        M.mul_(dy, db)
    ms = M.timer_stop() / reps
    ph, cnt = M.profile_read(5)
    per = [p / reps for p in ph]
    byts = [56, 96, 256, 96, 64]
    print("grid %d x %d x %d apply %.3f ms (%.0f GB/s on 568N)  " % (n, n, l, ms, 568 * N / ms / 1e6) +
          "  ".join("P%d %.3f ms %.0f GB/s" % (i + 1, per[i], byts[i] * N / per[i] / 1e6) for i in range(5)), flush=True)
    M.destroy(); db.free(); dy.free()
