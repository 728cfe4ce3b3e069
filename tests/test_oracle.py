"""CPU tests pinning the oracle (the reference has no golden vectors).

Known-answer substitutes:
  * dense Green matrix (buildConvMatrix, FastConvolution.jl:497-513) == trapezoidal FFT apply;
  * Greengard-Vico apply ~ trapezoidal apply to quadrature accuracy (sanity only);
  * pruned 4-way split identity used by the GPU kernels == the literal padded FFT;
  * As structural invariants; As*G far-field suppression;
  * GMRES restatement == exact minimal-residual Krylov solution (dense least squares);
  * committed golden fixtures reproduce.
"""
import os

import numpy as np
import pytest
import scipy.fft as sfft
import scipy.sparse as sp

from oracle import ls_oracle as O
from oracle.gmres_is import gmres, solve_least_squares

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def small_trap():
    h = 1 / 20
    n = 21
    x = -0.5 + h * np.arange(n)
    k = 1 / h
    M = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, "trapezoidal")
    X, Y = O.grid2d(x, x)
    return n, h, k, x, X, Y, M


def test_trapezoidal_apply_equals_dense_matrix(small_trap):
    n, h, k, x, X, Y, M = small_trap
    G = O.buildConvMatrix(k, X, Y, 1 - 0.892j, h)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    assert _rel(O.fastconvolution(M, b), b + k ** 2 * (G @ (M.nu * b))) < 1e-14
    # FFTconvolution (trapezoidal branch multiplies by nu, Q2)
    assert _rel(O.FFTconvolution(M, b), G @ (M.nu * b)) < 1e-13
    assert M.size(1) == n * n and M.size() == ((n * n,), (n * n,)) and M.eltype() == np.complex128
    Y_ = np.zeros(n * n, complex)
    M.mul_(Y_, b)
    assert np.array_equal(Y_, M * b)


def test_gv_close_to_trapezoidal_on_plane_wave(small_trap):
    n, h, k, x, X, Y, M = small_trap
    Mg = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, "Greengard_Vico")
    u = np.exp(1j * k * X)
    assert _rel(O.fastconvolution(Mg, u), O.fastconvolution(M, u)) < 2e-3
    # Q2: GV FFTconvolution applies no nu
    b = np.random.default_rng(1).standard_normal(n * n) + 0j
    assert _rel(b + k ** 2 * O.FFTconvolution(Mg, Mg.nu * b), O.fastconvolution(Mg, b)) < 1e-14


def test_gv_spectrum_symmetry_and_shift_folding():
    n = 16
    G = O.gv_spectrum_2d(n, n, 1 / n, 2 * np.pi * n / 10)
    assert np.array_equal(G[1:, :], G[:0:-1, :]) and np.array_equal(G[:, 1:], G[:, :0:-1])    # even in kx, ky
    assert np.array_equal(G, G.T)
    # fftshift/ifftshift around the product == multiplying by the pre-rolled spectrum (GPU create-time fold)
    rng = np.random.default_rng(2)
    B = rng.standard_normal((4 * n, 4 * n)) + 1j * rng.standard_normal((4 * n, 4 * n))
    F = np.fft.fft2(B)
    lit = np.fft.ifft2(np.fft.ifftshift(G * np.fft.fftshift(F)))
    fold = np.fft.ifft2(np.roll(G, (2 * n, 2 * n), axis=(0, 1)) * F)
    assert np.array_equal(lit, fold)


@pytest.mark.parametrize("n", [8, 32])
def test_pruned_split_identity(n):
    """X[4q+r] = FFT_n(x w_{4n}^{rj})[q] and the cropped inverse - the algebra of csrc/line_kernels.cuh."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    X = np.fft.fft(np.concatenate([x, np.zeros(3 * n)]))
    j = np.arange(n)
    for r in range(4):
        assert np.allclose(X[r::4], np.fft.fft(x * np.exp(-2j * np.pi * r * j / (4 * n))), atol=1e-12)
    Y = rng.standard_normal(4 * n) + 1j * rng.standard_normal(4 * n)
    y = np.fft.ifft(Y)[:n]
    acc = sum(np.exp(2j * np.pi * r * j / (4 * n)) * np.fft.ifft(Y[r::4]) for r in range(4)) / 4
    assert np.allclose(y, acc, atol=1e-13)


def test_3d_apply_against_direct_sum():
    """FastM3D `*` reproduces the direct quadrature sum of +exp(ikr)/(4 pi r) on a smooth source."""
    n = 16
    h = 1 / n
    x = -0.5 + h * np.arange(n)
    k = 1 / h / 4
    M = O.buildFastConvolution3D(x, x, x, h, k, O.nu_gaussian_3d)
    X, Y, Z = O.grid3d(x, x, x)
    f = np.exp(-60 * (X ** 2 + Y ** 2 + Z ** 2)).astype(complex)
    conv = O.FFTconvolution3D(M, f)
    idx = [0, 15 + n * 2 + n * n * 13, 1 + n * 14 + n * n * 0]       # far from the source: punctured sum is spectrally accurate
    for i in idx:
        r = np.sqrt((X - X[i]) ** 2 + (Y - Y[i]) ** 2 + (Z - Z[i]) ** 2)
        r[i] = 1.0
        g = np.exp(1j * k * r) / (4 * np.pi * r) * h ** 3
        g[i] = 0.0                       # punctured sum: O(h^2) accurate for a smooth density
        assert abs(conv[i] - g @ f) < 1e-3 * abs(conv[i])
    b = np.random.default_rng(3).standard_normal(n ** 3) + 0j
    assert _rel(M * b, b + k ** 2 * O.FFTconvolution3D(M, M.nu * b)) < 1e-15


def test_sparsifier_structure(small_trap):
    n, h, k, x, X, Y, M = small_trap
    D0 = 1 - 0.892j
    As = O.buildSparseA(k, X, Y, D0, n, n)
    assert As.nnz == 9 * (n - 2) ** 2 + 12 * (n - 2) + 12 * (n - 2) + 16
    counts = np.diff(As.tocsr().indptr)
    assert sorted(set(counts)) == [4, 6, 9]
    coo = As.tocoo()
    assert np.abs(coo.row - coo.col).max() == n + 1                       # half bandwidth
    # sparsification: |As G| away from the stencil is much smaller than inside it
    G = O.buildConvMatrix(k, X, Y, D0, h)
    AG = As @ G
    c = n * (n // 2) + n // 2
    row = np.abs(np.asarray(AG[c, :]).ravel())
    near = np.zeros(n * n, bool)
    near[[c + d for d in (-n - 1, -n, -n + 1, -1, 0, 1, n - 1, n, n + 1)]] = True
    assert row[~near].max() < 0.2 * row[near].max()
    # CSC column-scatter loop == scipy
    cp, rv, nz = O.julia_csc_arrays(As)
    b = np.random.default_rng(4).standard_normal(n * n) + 0j
    assert np.allclose(O.csc_matvec_loops(cp, rv, nz, b, n * n), As @ b, atol=1e-15)
    with pytest.raises(AssertionError):
        O.entriesSparseA(k, X[:400], Y[:400], D0, 20, 20)              # odd-N assert kept (SparsifyingMatrix2D.jl:7)


def test_gmres_is_minimal_residual(small_trap):
    n, h, k, x, X, Y, M = small_trap
    Mg = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, "Greengard_Vico")
    N = n * n
    A = np.column_stack([O.fastconvolution(Mg, e) for e in np.eye(N, dtype=complex)])
    rhs = -k ** 2 * O.FFTconvolution(Mg, Mg.nu * np.exp(1j * k * X))
    x0 = np.zeros(N, complex)
    xs, hist, conv, mv = gmres(x0, lambda v: A @ v, rhs, restart=20, reltol=1e-12)
    assert conv and mv == len(hist)
    # iteration j residual == min over the Krylov space K_j(A, rhs) (first cycle only)
    Kry = [rhs]
    for j in range(1, min(len(hist), 8) + 1):
        Kb = np.column_stack(Kry)
        c, *_ = np.linalg.lstsq(A @ Kb, rhs, rcond=None)
        assert abs(np.linalg.norm(rhs - A @ (Kb @ c)) - hist[j - 1]) < 1e-9 * np.linalg.norm(rhs)
        Kry.append(A @ Kry[-1])
    assert _rel(A @ xs, rhs) < 1e-10
    # restarts + left preconditioner: residuals are preconditioned residual norms
    D0 = 1 - 0.892j
    cache = O.entriesSparseA(k, X, Y, D0, n, n)
    As = O.buildSparseA(k, X, Y, D0, n, n, _cache=cache)
    AG = O.buildSparseAG(k, X, Y, D0, n, n, _cache=cache)
    P = O.SparsifyingPreconditioner(As + k ** 2 * (AG @ sp.diags(Mg.nu)), As)
    x1 = np.zeros(N, complex)
    x1, hist1, conv1, mv1 = gmres(x1, lambda v: A @ v, rhs, Pl_ldiv=P.solve, restart=3, reltol=1e-10)
    assert conv1 and len(hist1) < len(hist)
    assert _rel(x1, xs) < 1e-7
    true_prec_res = np.linalg.norm(P.solve(rhs - A @ x1))
    assert true_prec_res < 2e-10 * np.linalg.norm(P.solve(rhs)) * 10


def test_least_squares_givens():
    rng = np.random.default_rng(6)
    k = 7
    H = np.triu(rng.standard_normal((k, k - 1)) + 1j * rng.standard_normal((k, k - 1)), -1)
    y = solve_least_squares(H, 2.5, k)
    rhs = np.zeros(k, complex); rhs[0] = 2.5
    yr, *_ = np.linalg.lstsq(H, rhs, rcond=None)
    assert np.allclose(y, yr, atol=1e-12)


def test_golden_fixtures_reproduce():
    g = np.load(os.path.join(GOLD, "apply2d_n64.npz"))
    n = int(g["n"])
    x, h, k, M = O.pow2_problem_2d(n)
    b = np.random.default_rng(int(g["seed"])).standard_normal(n * n) + 1j * np.random.default_rng(int(g["seed"])).standard_normal(n * n)
    rng = np.random.default_rng(int(g["seed"]))
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    assert _rel(O.fastconvolution(M, b), g["y_fastconvolution"]) < 1e-14
    assert _rel(O.FFTconvolution(M, b), g["y_FFTconvolution"]) < 1e-14
    gg = np.load(os.path.join(GOLD, "gmres2d_n64.npz"))
    assert gg["hist_precond"].shape[0] < gg["hist_plain"].shape[0]


def test_sparsifier_3d_restatement():
    """SparsifyingMatrix3D.jl:963-1135, 1136-1408, 1410-1653, 1659-1917 restated (example3D.jl:56-61): structure of As,
    the literal sampleG3D against the Toeplitz shortcut, far-field suppression and the preconditioner's effect."""
    from oracle.gmres_is import gmres
    n, l = 8, 10
    (x, z), h, k, M, As, Msp, P = O.example_problem_3d(n, l)
    N = n * n * l
    X, Y, Z = O.grid3d(x, x, z)
    # 27 boundary classes: interior 27 entries, faces 18, edges 12, corners 8
    assert As.shape == (N, N)
    assert As.nnz == (n - 2) ** 2 * (l - 2) * 27 + (2 * (n - 2) ** 2 + 4 * (n - 2) * (l - 2)) * 18 \
        + (8 * (n - 2) + 4 * (l - 2)) * 12 + 64
    coo = As.tocoo()
    assert np.max(np.abs(coo.row - coo.col)) == n * n + n + 1                 # half bandwidth nm + n + 1
    csr = As.tocsr()
    classes = {(tuple(csr.indices[csr.indptr[r]:csr.indptr[r + 1]] - r), csr.data[csr.indptr[r]:csr.indptr[r + 1]].tobytes())
               for r in range(N)}
    assert len(classes) == 27
    # sampleG3D: applies of unit vectors (FastConvolution3D.jl:136-160) == shifted copies of the spatial kernel
    ind = np.array([1, n, n * n + 3, N // 2, N])
    lit = O.sampleG3D(k, X, Y, Z, ind, M, toeplitz=False)
    assert np.abs(lit - O.sampleG3D(k, X, Y, Z, ind, M)).max() <= 1e-14 * np.abs(lit).max()
    # each stencil is a unit vector (last left singular vector) ...
    Indices, Values = O.entriesSparseA3D(k, X, Y, Z, M, n, n, l)
    assert [len(v) for v in Values] == [27] + [18] * 6 + [12] * 12 + [8] * 8
    assert all(abs(np.linalg.norm(v) - 1) < 1e-12 for v in Values)
    # ... chosen so that As*G is (nearly) supported on the stencil: the truncated AG reproduces it
    G = O.sampleG3D(k, X, Y, Z, np.arange(1, N + 1), M)
    AsG = As @ G
    AG = O.buildSparseAG3DConv(k, X, Y, Z, M, n, n, l).toarray()
    assert np.allclose(AG[AG != 0], AsG[AG != 0], rtol=1e-10, atol=1e-14)       # same numbers on the stencil
    assert np.linalg.norm(AsG - AG) < 0.25 * np.linalg.norm(AsG)                 # little left outside
    # the preconditioned solve needs no more iterations and gives the same solution
    u_inc = np.exp(1j * k * X)
    rhs = -(M * u_inc - u_inc)
    u0, h0, c0, _ = gmres(np.zeros(N, complex), lambda v: M * v, rhs)
    u1, h1, c1, _ = gmres(np.zeros(N, complex), lambda v: M * v, rhs, Pl_ldiv=P.solve)
    assert c0 and c1 and len(h1) <= len(h0)
    assert np.linalg.norm(u1 - u0) <= 1e-6 * np.linalg.norm(u0)


def test_compact_padding_identity_and_mirror_symmetry():
    """The device path's central algebra, checked on the CPU: because the input lives on [0, n) and the output is
    cropped to [0, n), the 4x-padded apply only uses g = ifft2(ifftshift(GFFT)) at the lags (-n, n); wrapping those
    onto a 2n grid and transforming back gives a 2x-padded apply that is the same operator.  The
    wrapped kernel is even in every coordinate, so its spectrum is mirror symmetric."""
    n = 24
    x, h, k, M = O.pow2_problem_2d(n, ppw=9.3)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    ref = O.FFTconvolution(M, b)                                   # literal: pad to 4n, full FFTs, crop
    g4 = sfft.ifft2(sfft.ifftshift(M.GFFT))                        # spatial kernel on the 4n grid
    lag = np.arange(-n, n)                                         # lags kept, placed at lag mod 2n
    g2 = np.zeros((2 * n, 2 * n), complex)
    g2[np.ix_(lag % (2 * n), lag % (2 * n))] = g4[np.ix_(lag % (4 * n), lag % (4 * n))]
    G2 = sfft.fft2(g2)
    B = np.zeros((2 * n, 2 * n), complex)
    B[:n, :n] = b.reshape((n, n), order="F")
    y = sfft.ifft2(G2 * sfft.fft2(B))[:n, :n].reshape(-1, order="F")
    assert np.linalg.norm(y - ref) <= 1e-13 * np.linalg.norm(ref)
    # lag -n is never used by the cropped apply (|i - j| <= n - 1): dropping it makes the kernel exactly even
    g2e = g2.copy()
    g2e[n, :] = 0
    g2e[:, n] = 0
    y2 = sfft.ifft2(sfft.fft2(g2e) * sfft.fft2(B))[:n, :n].reshape(-1, order="F")
    assert np.linalg.norm(y2 - ref) <= 1e-13 * np.linalg.norm(ref)
    G2e = sfft.fft2(g2e)
    idx = (-np.arange(2 * n)) % (2 * n)
    scale = np.abs(G2e).max()
    assert np.abs(G2e - G2e[idx, :]).max() <= 1e-12 * scale and np.abs(G2e - G2e[:, idx]).max() <= 1e-12 * scale
