"""Checks that pin the oracle and the product's host logic WITHOUT going through the code they share
("parity unpinned" items): none of these uses the oracle's FFT path or the sibling sparsifier port.

  * the 2-D Greengard-Vico apply against a long-double direct evaluation: the spatial kernel is the inverse DFT of
    the given spectrum summed term by term in extended precision, the apply is the O(N^2) convolution sum
    (FastConvolution.jl:84-106 says: pad, fftshift(fft), multiply, ifft(ifftshift), crop - every one of those steps
    is replaced by its definition here);
  * the product's sparsifier (fast_solver_lippmann_schwinger_b200.sparsifier, rows sampled through applies, QR + small
    SVD) against null vectors computed from the dense Green matrix of buildConvMatrix (FastConvolution.jl:497-513)
    with a plain dense SVD and an independently written stencil enumeration.
"""
import numpy as np
import pytest

from oracle import ls_oracle as O


def _gv_kernel_longdouble(GFFT, lags):
    """g[dx, dy] = 1/(ne me) sum_{kx,ky} GFFT_centred[kx, ky] exp(+2 pi i (kx dx/ne + ky dy/me)), kx = -ne/2..ne/2-1,
    in long double (the definition of ifft(ifftshift(.)) with the centred ordering of FastConvolution.jl:220-226)."""
    ne, me = GFFT.shape
    kx = np.arange(-ne // 2, ne // 2, dtype=np.longdouble)
    ky = np.arange(-me // 2, me // 2, dtype=np.longdouble)
    G = GFFT.astype(np.clongdouble)
    two_pi = 2 * np.pi.__class__(np.pi) if False else np.longdouble(2) * np.arccos(np.longdouble(-1))
    out = {}
    for dx in lags:
        ex = np.exp(1j * (two_pi * kx * np.longdouble(dx) / ne))          # (ne,)
        row = ex @ G                                                       # sum over kx -> (me,)
        for dy in lags:
            ey = np.exp(1j * (two_pi * ky * np.longdouble(dy) / me))
            out[(dx, dy)] = (row @ ey) / (np.longdouble(ne) * me)
    return out


@pytest.mark.parametrize("n", [8, 12])
def test_gv_apply_against_longdouble_direct_sum(n):
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    k = 2 * np.pi / (7.3 * h)
    M = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    rng = np.random.default_rng(n)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    g = _gv_kernel_longdouble(M.GFFT, range(-(n - 1), n))
    f = (M.nu * b).reshape(n, n, order="F").astype(np.clongdouble)
    conv = np.zeros((n, n), dtype=np.clongdouble)
    for ix in range(n):
        for iy in range(n):
            acc = np.clongdouble(0)
            for jx in range(n):
                for jy in range(n):
                    acc += g[(ix - jx, iy - jy)] * f[jx, jy]
            conv[ix, iy] = acc
    y_direct = b + (k ** 2) * conv.reshape(-1, order="F").astype(np.complex128)
    y = O.fastconvolution(M, b)
    assert np.linalg.norm(y - y_direct) / np.linalg.norm(y_direct) < 1e-13
    # FFTconvolution (GV branch: no nu, Q2) through the same kernel
    fb = b.reshape(n, n, order="F").astype(np.clongdouble)
    c0 = sum(g[(0 - jx, 3 - jy)] * fb[jx, jy] for jx in range(n) for jy in range(n))
    assert abs(O.FFTconvolution(M, b)[0 + n * 3] - complex(c0)) < 1e-13 * abs(complex(c0))


def _null_vector_dense(G, stencil):
    """Last left singular vector of the far-field block G[stencil, far] by a plain dense SVD (what
    entriesSparseA does with U[:, end]', SparsifyingMatrix2D.jl:5-102)."""
    far = np.ones(G.shape[1], dtype=bool)
    far[stencil] = False
    U, s, Vh = np.linalg.svd(G[np.ix_(stencil, np.flatnonzero(far))], full_matrices=False)
    return np.conj(U[:, -1])


def test_product_sparsifier_against_dense_green_matrix():
    """As rows of the product's buildSparseAConv at the class representatives == dense null vectors (up to a phase)."""
    from fast_solver_lippmann_schwinger_b200 import sparsifier as S
    n = 15
    h = 1.0 / (n - 1)
    x = -0.5 + h * np.arange(n)
    k = 1.0 / h
    D0 = 1 - 0.892j
    X, Y = O.grid2d(x, x)
    G = O.buildConvMatrix(k, X, Y, D0, h)                     # dense N x N Green matrix, FastConvolution.jl:497-513

    def apply(M, e):                                           # FFTconvolution(fastconv, e) for a unit contrast == G e
        return G @ e

    As = S.buildSparseAConv(k, X, Y, object(), n, n, apply=apply).tocsr()
    N = n * n
    # independent enumeration: a grid point (i, j) (0-based) and its in-grid neighbours, x fastest
    def stencil(i, j):
        return [ii + n * jj for jj in (j - 1, j, j + 1) for ii in (i - 1, i, i + 1) if 0 <= ii < n and 0 <= jj < n]

    c = n // 2
    m = n
    # the class representatives, straight from SparsifyingMatrix2D.jl:119,131,140,149,158,166-169 (1-based there); note the
    # upstream quirk that the x = xmax and y = ymax edges sit one line off the middle (n(m-1)/2 and N - (n+1)/2)
    one_based = {"interior": n * (m - 1) // 2 + (n + 1) // 2, "x-lo edge": n * (m - 1) // 2 + 1, "x-hi edge": n * (m - 1) // 2,
                 "y-lo edge": (n + 1) // 2, "y-hi edge": N - (n + 1) // 2,
                 "corner 00": 1, "corner n0": n, "corner 0n": N - n + 1, "corner nn": N}
    reps = {name: ((v - 1) % n, (v - 1) // n) for name, v in one_based.items()}
    assert reps["interior"] == (c, c) and reps["x-hi edge"] == (n - 1, c - 1) and reps["y-hi edge"] == (c - 1, n - 1)
    for name, (i, j) in reps.items():
        st = stencil(i, j)
        v = _null_vector_dense(G, st)
        r = i + n * j
        cols = As.indices[As.indptr[r]:As.indptr[r + 1]]
        vals = As.data[As.indptr[r]:As.indptr[r + 1]]
        assert sorted(cols) == sorted(st), name
        row = np.zeros(N, complex)
        row[cols] = vals
        a = row[st]
        ph = np.vdot(v, a) / abs(np.vdot(v, a))
        assert np.abs(a - ph * v).max() < 1e-8 * np.abs(v).max(), name
        # and it does sparsify: |(As G)[r, far]| is far below |(As G)[r, stencil]|
        ag = row @ G
        far = np.ones(N, bool); far[st] = False
        assert np.abs(ag[far]).max() < 0.25 * np.abs(ag[st]).max(), name
    # translation invariance: every interior row carries the interior representative's coefficients
    r0 = c + n * c
    base = As.data[As.indptr[r0]:As.indptr[r0 + 1]]
    for (i, j) in ((1, 1), (n - 2, 3), (5, n - 2)):
        r = i + n * j
        assert np.array_equal(As.indices[As.indptr[r]:As.indptr[r + 1]] - r, As.indices[As.indptr[r0]:As.indptr[r0 + 1]] - r0)
        assert np.array_equal(As.data[As.indptr[r]:As.indptr[r + 1]], base)


def test_product_sparsifier_3d_against_dense_green_matrix():
    """3-D: every row of the product's buildSparseA3DConv against null vectors of the dense Green matrix, with the boundary
    classes, their representatives and the stencils enumerated geometrically here (not through the class tables the
    product and the oracle's sparsifier section share).  Representatives as in SparsifyingMatrix3D.jl:1166-1341:
    coordinate 1 on a low face, n on a high face, round(n/2) otherwise (changeInd3D(...) calls at :1166, :1178, :1189,
    :1199, :1208, :1217, :1225, :1233-1244, :1334-1341); stencil order = Ind_relative[...][:], x fastest."""
    from fast_solver_lippmann_schwinger_b200 import sparsifier as S
    n, m, l = 6, 6, 5                                        # FFTconvolution 3-D pads (ne, ne, le): n == m (Q3)
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    z = -0.5 + h * np.arange(l)
    k = 2 * np.pi / (7.3 * h)
    Mo = O.buildFastConvolution3D(x, x, z, h, k, O.nu_gaussian_3d)
    N = n * m * l
    G = np.empty((N, N), complex)                            # dense matrix of the convolution: column c = apply(e_c)
    e = np.zeros(N, complex)
    for c in range(N):
        e[c] = 1.0
        G[:, c] = O.FFTconvolution3D(Mo, e)
        e[c] = 0.0
    X, Y, Z = O.grid3d(x, x, z)
    As = S.buildSparseA3DConv(k, X, Y, Z, object(), n, m, l, apply=lambda M, v: G @ v).tocsr()

    def rnd_half(v):                                         # Julia's round(Integer, v/2): ties to even
        return int(np.round(v / 2))

    def rep_coord(c, size):                                  # 1-based coordinate of the class representative along one axis
        return 1 if c == 1 else (size if c == size else rnd_half(size))

    def neighbours(i, j, p):                                 # 1-based coordinates -> 0-based linear indices, x fastest
        return [(ii - 1) + n * (jj - 1) + n * m * (pp - 1)
                for pp in (p - 1, p, p + 1) for jj in (j - 1, j, j + 1) for ii in (i - 1, i, i + 1)
                if 1 <= ii <= n and 1 <= jj <= m and 1 <= pp <= l]

    cache = {}
    sizes = set()
    for p in range(1, l + 1):
        for j in range(1, m + 1):
            for i in range(1, n + 1):
                rep = (rep_coord(i, n), rep_coord(j, m), rep_coord(p, l))
                if rep not in cache:
                    cache[rep] = _null_vector_dense(G, neighbours(*rep))
                v = cache[rep]
                st = neighbours(i, j, p)
                assert len(st) == v.size
                sizes.add(v.size)
                r = (i - 1) + n * (j - 1) + n * m * (p - 1)
                cols = As.indices[As.indptr[r]:As.indptr[r + 1]]
                vals = As.data[As.indptr[r]:As.indptr[r + 1]]
                assert sorted(cols) == st, (i, j, p)
                a = vals[np.argsort(cols)]
                ph = np.vdot(v, a) / abs(np.vdot(v, a))
                assert np.abs(a - ph * v).max() < 1e-7 * np.abs(v).max(), (i, j, p)
    assert len(cache) == 27 and sizes == {27, 18, 12, 8}
