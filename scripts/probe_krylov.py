"""Probe: GMRES time-to-tolerance (unpreconditioned), SpMV and MGS bandwidth on one B200."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import fast_solver_lippmann_schwinger_b200 as ls
from fast_solver_lippmann_schwinger_b200.problems import gv_problem_2d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nu, gfft, k, h = gv_problem_2d(n)
M = ls.FastM(gfft, nu, 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico"); del gfft
N = n * n
x = -0.5 + h * np.arange(n)
X = np.repeat(x[:, None], n, axis=1).reshape(-1, order="F")
u_inc = np.exp(1j * k * X)
rhs = -(M * u_inc - u_inc)
db = ls.DeviceBuffer.from_host(rhs); dx = ls.DeviceBuffer.from_host(np.zeros(N, complex))
ws = ls.KrylovWorkspace(N)
t0 = time.perf_counter()
_, hist = ls.gmres_(dx, M, db, reltol=1e-8, maxiter=maxit, log=True, workspace=ws)
dt = time.perf_counter() - t0
print("GMRES n=%d iters=%d conv=%s time=%.3fs (%.3f ms/iter) res0=%.3e last=%.3e" % (n, hist.iters, hist.isconverged, dt, 1e3 * dt / max(hist.iters, 1), hist["resnorm"][0], hist["resnorm"][-1]))
# SpMV: synthetic 9-point stencil matrix in Julia CSC layout
idx = np.arange(N).reshape(n, n, order="F")
rows, cols = [], []
for di in (-1, 0, 1):
    for dj in (-1, 0, 1):
        src = idx[max(0, -di):n - max(0, di), max(0, -dj):n - max(0, dj)]
        dst = idx[max(0, di):n - max(0, -di), max(0, dj):n - max(0, -dj)]
        rows.append(src.ravel()); cols.append(dst.ravel())
rows = np.concatenate(rows); cols = np.concatenate(cols)
rng = np.random.default_rng(0)
A = sp.csc_matrix((rng.standard_normal(rows.size) + 1j * rng.standard_normal(rows.size), (rows, cols)), shape=(N, N))
dy = ls.DeviceBuffer(16 * N)
alg = A.nnz * 20 + 4 * (N + 1) + 32 * N
# translation-invariant values (what buildSparseA produces): class structure -> stencil path
st = rng.standard_normal(9) + 1j * rng.standard_normal(9)
vals = np.concatenate([np.full(r.size, st[i]) for i, r in enumerate(np.split(rows, np.cumsum([((n - abs(di)) * (n - abs(dj))) for di in (-1, 0, 1) for dj in (-1, 0, 1)])[:-1]))])
A2 = sp.csc_matrix((vals, (rows, cols)), shape=(N, N))
for name, mat in (("random values", A), ("stencil values", A2)):
    G = ls.GPUSparseMatrixCSC(mat)
    for _ in range(3): G.mv(db, dy)
    G.sync(); G.timer_start()
    for _ in range(20): G.mv(db, dy)
    ms = G.timer_stop() / 20
    print("SpMV %s format=%s classes=%d nnz=%d %.4f ms  CSR-accounting %.0f GB/s (%.1f%% of 6551)" % (name, G.format, G.nclasses, mat.nnz, ms, alg / ms / 1e6, alg / ms / 1e6 / 65.51))
# MGS step k=10, 20
for kk in (10, 20):
    V = ls.DeviceBuffer(16 * N * (kk + 1))
    ls.lib().ls_memcpy_h2d  # noqa
    w = ls.DeviceBuffer.from_host(rhs)
    ws.mgs_step(V, N, kk, w)
    t0 = time.perf_counter()
    for _ in range(10): ws.mgs_step(V, N, kk, w)
    ms = (time.perf_counter() - t0) / 10 * 1e3
    alg = (64 * kk + 48) * N
    print("MGS k=%d %.4f ms alg(64k+48)N %.0f GB/s (%.1f%%)" % (kk, ms, alg / ms / 1e6, alg / ms / 1e6 / 65.51))
