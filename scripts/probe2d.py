"""Quick timing probe of the 2-D apply (device-resident), not the bench."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fast_solver_lippmann_schwinger_b200 as ls
from fast_solver_lippmann_schwinger_b200.problems import gv_problem_2d

sizes = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048]
for n in sizes:
    t0 = time.time()
    nu, gfft, k, h = gv_problem_2d(n)
    t1 = time.time()
    M = ls.FastM(gfft, nu, 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico")
    del gfft
    t2 = time.time()
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    db = ls.DeviceBuffer.from_host(b); dy = ls.DeviceBuffer(b.nbytes)
    for _ in range(3):
        M.mul_(dy, db)
    M.sync()
    reps = 20
    M.timer_start()
    for _ in range(reps):
        M.mul_(dy, db)
    ms = M.timer_stop() / reps
    N = n * n
    print("n=%d setup %.1fs create %.1fs  apply %.4f ms  -> %.1f applies/s, alg %.1f GB/s (%.1f%% of 6551)" % (
        n, t1 - t0, t2 - t1, ms, 1e3 / ms, 568 * N / ms / 1e6, 568 * N / ms / 1e6 / 6551 * 100), flush=True)
    M.destroy()
