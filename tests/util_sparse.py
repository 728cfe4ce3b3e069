"""Synthetic 27-point matrices with the sparsity structure of the 3-D sparsifier (SparsifyingMatrix3D.jl:1147-1158,
1410-1653): row (i, j, p) of the n x m x l grid (x fastest) couples to its <= 27 grid neighbours."""
import numpy as np
import scipy.sparse as sp


def stencil27(n, m, l, seed=0, classes=True):
    """classes=True: one coefficient vector per boundary class (interior, faces, edges, corners - 27 classes), as the
    reference builds it; classes=False: independent random values (no class structure -> CSR path)."""
    rng = np.random.default_rng(seed)
    I, J, P = np.meshgrid(np.arange(n), np.arange(m), np.arange(l), indexing="ij")
    I, J, P = (a.reshape(-1, order="F") for a in (I, J, P))
    row = I + n * (J + m * P)
    bx = np.where(I == 0, 0, np.where(I == n - 1, 2, 1))
    by = np.where(J == 0, 0, np.where(J == m - 1, 2, 1))
    bz = np.where(P == 0, 0, np.where(P == l - 1, 2, 1))
    cls = bx + 3 * (by + 3 * bz)
    coef = rng.standard_normal((27, 27)) + 1j * rng.standard_normal((27, 27))
    rows, cols, vals = [], [], []
    q = 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = (I + dx >= 0) & (I + dx < n) & (J + dy >= 0) & (J + dy < m) & (P + dz >= 0) & (P + dz < l)
                rows.append(row[ok])
                cols.append(row[ok] + dx + n * (dy + m * dz))
                vals.append(coef[cls[ok], q] if classes else rng.standard_normal(ok.sum()) + 1j * rng.standard_normal(ok.sum()))
                q += 1
    N = n * m * l
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    A.sort_indices()
    return A
