"""Time one orthogonalisation sweep (MGS / CGS / DGKS kernels) at 2048^2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fast_solver_lippmann_schwinger_b200 as ls
from fast_solver_lippmann_schwinger_b200._lib import check, lib
N = 2048 * 2048
rng = np.random.default_rng(0)
ws = ls.KrylovWorkspace(N)
w0 = rng.standard_normal(N) + 1j * rng.standard_normal(N)
for kk in (5, 10, 20):
    V = ls.DeviceBuffer.from_host((rng.standard_normal(N * kk) + 1j * rng.standard_normal(N * kk)) / np.sqrt(N))
    for meth in (0, 1, 2):
        check(lib().ls_krylov_set_orth(ws.handle, meth))
        w = ls.DeviceBuffer.from_host(w0)
        ws.mgs_step(V, N, kk, w)
        t0 = time.perf_counter()
        for _ in range(10):
            ws.mgs_step(V, N, kk, w)
        ms = (time.perf_counter() - t0) / 10 * 1e3
        print("k=%d orth=%d %.4f ms" % (kk, meth, ms), flush=True)
    V.free()
