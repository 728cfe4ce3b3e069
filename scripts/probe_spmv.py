import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
import fast_solver_lippmann_schwinger_b200 as ls
n = 2048; N = n * n
rng = np.random.default_rng(7)
idx = np.arange(N).reshape(n, n, order="F")
rows, cols = [], []
for di in (-1, 0, 1):
    for dj in (-1, 0, 1):
        rows.append(idx[max(0, -di):n - max(0, di), max(0, -dj):n - max(0, dj)].ravel())
        cols.append(idx[max(0, di):n - max(0, -di), max(0, dj):n - max(0, -dj)].ravel())
rows = np.concatenate(rows); cols = np.concatenate(cols)
A = sp.csc_matrix((rng.standard_normal(rows.size) + 1j * rng.standard_normal(rows.size), (rows, cols)), shape=(N, N))
x = ls.DeviceBuffer.from_host(rng.standard_normal(N) + 1j * rng.standard_normal(N)); y = ls.DeviceBuffer(16 * N)
alg = A.nnz * 20 + 4 * (N + 1) + 32 * N
for lanes in (1, 2, 4, 8):
    os.environ["LS_SPM_LANES"] = str(lanes)
    G = ls.GPUSparseMatrixCSC(A)
    for _ in range(3): G.mv(x, y)
    G.sync(); G.timer_start()
    for _ in range(20): G.mv(x, y)
    ms = G.timer_stop() / 20
    print("lanes", lanes, "%.4f ms" % ms, "%.0f GB/s %.1f%%" % (alg / ms / 1e6, alg / ms / 1e6 / 65.51))
