"""Quick timing probe of the 3-D apply (device-resident), not the bench."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fast_solver_lippmann_schwinger_b200 as ls
from fast_solver_lippmann_schwinger_b200.problems import nu_gaussian_3d_grid

sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
for n in sizes:
    h = 1.0 / n; k = 2 * np.pi / (10 * h)
    t0 = time.time()
    nu = nu_gaussian_3d_grid(n)
    M = ls.FastM3D(None, nu, 4 * n, 4 * n, 4 * n, n, n, n, k, L=1.8 * n * h, Lp=4.0 * n * h)
    t1 = time.time()
    N = n ** 3
    rng = np.random.default_rng(1234)
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
    print("n=%d create %.1fs  apply %.3f ms -> %.1f applies/s, alg %.0f GB/s (%.1f%% of 6551) phases(ms) %s" % (
        n, t1 - t0, ms, 1e3 / ms, 2360 * N / ms / 1e6, 2360 * N / ms / 1e6 / 6551 * 100,
        ["%.3f" % (p / reps) for p in ph]) + " chunks=%s" % os.environ.get("LS_OP3D_CHUNKS", "1"), flush=True)
    M.destroy()
