"""GPU parity of the sparsifying SpMV, the Arnoldi vector kernels and the GMRES driver.

Tolerances: SpMV / BLAS-1 1e-13 relative (summation order differs from the CPU loop only at
rounding level); GMRES residual histories agree with the oracle to 1e-8 relative over the first
50 iterations (BASELINE.json north_star).
"""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _sparsifier_problem(n):
    """GV operator + sparsifying preconditioner matrices from the oracle's restatement."""
    from oracle import ls_oracle as O
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    k = 1.0 / h                                   # kh = 1 as in examples/example.jl:31
    Mo = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    X, Y = O.grid2d(x, x)
    D0 = O.referenceValsTrapRule()[1][0]
    cache = O.entriesSparseA(k, X, Y, D0, n, n, strict=False)
    As = O.buildSparseA(k, X, Y, D0, n, n, strict=False, _cache=cache)
    AG = O.buildSparseAG(k, X, Y, D0, n, n, strict=False, _cache=cache)
    Msp = (As + k ** 2 * (AG @ sp.diags(Mo.nu))).tocsc()
    return Mo, As, Msp, X, k


def test_spmv_matches_csc_loop():
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    Mo, As, Msp, X, k = _sparsifier_problem(64)
    N = As.shape[0]
    rng = np.random.default_rng(5)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    A = ls.GPUSparseMatrixCSC(As)
    assert A.nnz == As.nnz == 9 * 62 * 62 + 24 * 62 + 16
    assert A.format == "stencil" and A.nclasses == 9          # the 9 boundary classes of buildSparseA
    y = A * x
    assert _rel(y, O.csc_matvec(As, x)) < 1e-13
    # the general CSR kernel on the same matrix (structure detection switched off)
    import os
    os.environ["LS_SPM_FORCE_CSR"] = "1"
    try:
        Acsr = ls.GPUSparseMatrixCSC(As)
    finally:
        del os.environ["LS_SPM_FORCE_CSR"]
    assert Acsr.format == "csr"
    assert _rel(Acsr * x, O.csc_matvec(As, x)) < 1e-13
    # Msp has a position-dependent contrast: no class structure, CSR path
    assert ls.GPUSparseMatrixCSC(Msp).format == "csr"
    assert _rel(ls.GPUSparseMatrixCSC(Msp) * x, Msp @ x) < 1e-13
    # Julia-layout arrays + the pure loop statement on a small slice of columns
    cp, rv, nz = O.julia_csc_arrays(As)
    A2 = ls.GPUSparseMatrixCSC((N, N, cp, rv, nz))
    assert np.array_equal(A2 * x, y)
    # cscmv! semantics: y <- alpha A x + beta y
    y0 = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    al, be = 0.7 - 0.2j, -1.3 + 0.5j
    y1 = y0.copy()
    ls.cscmv_("N", al, "GXXF", A, x, be, y1)
    assert _rel(y1, al * (As @ x) + be * y0) < 1e-13
    with pytest.raises(ValueError):
        ls.cscmv_("N", 1.0, "GXXF", A, x[:-1], 0.0, y1)


def test_spmv_ragged_and_empty_rows():
    import fast_solver_lippmann_schwinger_b200 as ls
    rng = np.random.default_rng(11)
    n = 1000
    A = sp.random(n, n, density=0.02, random_state=3, format="csc", dtype=np.float64)
    A = (A + 1j * sp.random(n, n, density=0.02, random_state=4, format="csc")).tocsc()
    A = A.tolil(); A[17, :] = 0; A[:, 5] = 0; A = A.tocsc(); A.eliminate_zeros()      # an empty row and an empty column
    dense_row = sp.csc_matrix((np.ones(n) * (1 + 2j), (np.full(n, 33), np.arange(n))), shape=(n, n))
    A = (A + dense_row).tocsc()                                                       # one long row (n entries)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    G = ls.GPUSparseMatrixCSC(A)
    assert _rel(G * x, A @ x) < 1e-13
    # rectangular
    B = sp.random(300, 500, density=0.05, random_state=8, format="csc").astype(np.complex128)
    xb = rng.standard_normal(500) + 1j * rng.standard_normal(500)
    assert _rel(ls.GPUSparseMatrixCSC(B) * xb, B @ xb) < 1e-13


def test_blas1_and_mgs():
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 300_007                     # ragged against the block size
    rng = np.random.default_rng(2)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    K = ls.KrylovWorkspace(n)
    dx, dy = ls.DeviceBuffer.from_host(x), ls.DeviceBuffer.from_host(y)
    d = K.dot(dx, dy)
    assert abs(d - np.vdot(x, y)) < 1e-12 * np.linalg.norm(x) * np.linalg.norm(y)
    assert d == K.dot(dx, dy)                              # bitwise reproducible
    assert abs(K.norm(dx) - np.linalg.norm(x)) < 1e-13 * np.linalg.norm(x)
    al = 0.3 - 2.0j
    K.axpy(al, dx, dy); K.sync()
    assert _rel(dy.to_host(), y + al * x) < 1e-15
    K.scal(al, dx); K.sync()
    assert _rel(dx.to_host(), al * x) < 1e-15
    # modified Gram-Schmidt against k orthonormal columns
    for k in (1, 2, 7, 20):
        Q, _ = np.linalg.qr(rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k)))
        V = np.asfortranarray(Q)
        w = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        dV = ls.DeviceBuffer.from_host(V.reshape(-1, order="F"))
        dw = ls.DeviceBuffer.from_host(w)
        h = K.mgs_step(dV, n, k, dw)
        wr = w.copy(); hr = np.zeros(k + 1, complex)
        for i in range(k):
            hr[i] = np.vdot(V[:, i], wr); wr = wr - hr[i] * V[:, i]
        hr[k] = np.linalg.norm(wr); wr /= hr[k]
        assert np.abs(h - hr).max() < 1e-12 * np.abs(hr).max()
        assert _rel(dw.to_host(), wr) < 1e-12


@pytest.mark.parametrize("precond", [False, True])
def test_gmres_history_matches_oracle(precond):
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres as gmres_oracle
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 64
    Mo, As, Msp, X, k = _sparsifier_problem(n)
    N = n * n
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega, quadRule="Greengard_Vico")
    u_inc = np.exp(1j * k * X)
    rhs = -k ** 2 * O.FFTconvolution(Mo, Mo.nu * u_inc)            # examples/example.jl:76-77
    if precond:
        Po = O.SparsifyingPreconditioner(Msp, As)
        Pg = ls.SparsifyingPreconditioner(Msp, As)
        pl_o = Po.solve
    else:
        Pg, pl_o = None, None
    xo = np.zeros(N, complex)
    xo, hist_o, conv_o, mv_o = gmres_oracle(xo, lambda v: O.fastconvolution(Mo, v), rhs, Pl_ldiv=pl_o, maxiter=60)
    xg = np.zeros(N, complex)
    xg, hg = ls.gmres_(xg, Mg, rhs, Pl=Pg, log=True, maxiter=60)
    assert hg.iters == len(hist_o) and hg.isconverged == conv_o and hg.mvps == mv_o
    m = min(50, len(hist_o))
    assert np.max(np.abs(hg["resnorm"][:m] - hist_o[:m]) / hist_o[:m]) < 1e-8
    assert _rel(xg, xo) < 1e-8
    # the solution solves the system
    res = np.linalg.norm(O.fastconvolution(Mo, xg) - rhs) / np.linalg.norm(rhs)
    assert res < 1e-6


def test_gmres_restart_and_device_vectors():
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres as gmres_oracle
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 64
    x, h, k, Mo = O.pow2_problem_2d(n, ppw=6.0, nu=lambda X, Y: 4.0 * O.nu_gaussian_2d(X, Y))   # harder: needs restarts
    N = n * n
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega, quadRule="Greengard_Vico")
    rng = np.random.default_rng(9)
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    xo = np.zeros(N, complex)
    xo, hist_o, conv_o, mv_o = gmres_oracle(xo, lambda v: O.fastconvolution(Mo, v), rhs, restart=5, maxiter=23, reltol=1e-10)
    db = ls.DeviceBuffer.from_host(rhs)
    dx = ls.DeviceBuffer.from_host(np.zeros(N, complex))
    _, hg = ls.gmres_(dx, Mg, db, restart=5, maxiter=23, reltol=1e-10, log=True)
    assert hg.iters == len(hist_o) == 23 and hg.mvps == mv_o
    assert np.max(np.abs(hg["resnorm"] - hist_o) / hist_o) < 1e-8
    assert _rel(dx.to_host(), xo) < 1e-8


def test_plasma_contrast_preconditioned_solve():
    """Config 3 (tests/plasma_example.jl) at reduced size: discontinuous plasma contrast, Greengard_Vico
    operator, sparsifying preconditioner, rhs = -(A u_inc - u_inc) (plasma_example.jl:160-161)."""
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres as gmres_oracle
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 128
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    k = 1.0 / h
    Mo = O.buildFastConvolution(x, x, h, k, O.nu_plasma_2d, quadRule="Greengard_Vico")
    assert np.unique(Mo.nu).size > 100 and (Mo.nu == 0).sum() > 100          # genuinely discontinuous profile
    X, Y = O.grid2d(x, x)
    D0 = O.referenceValsTrapRule()[1][0]
    cache = O.entriesSparseA(k, X, Y, D0, n, n, strict=False)
    As = O.buildSparseA(k, X, Y, D0, n, n, strict=False, _cache=cache)
    AG = O.buildSparseAG(k, X, Y, D0, n, n, strict=False, _cache=cache)
    Msp = (As + k ** 2 * (AG @ sp.diags(Mo.nu))).tocsc()
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, n, k, quadRule="Greengard_Vico")
    u_inc = np.exp(1j * k * X)
    rhs_o = -(O.fastconvolution(Mo, u_inc) - u_inc)
    rhs_g = -(Mg * u_inc - u_inc)
    assert _rel(rhs_g, rhs_o) < 1e-12
    Po, Pg = O.SparsifyingPreconditioner(Msp, As), ls.SparsifyingPreconditioner(Msp, As)
    xo = np.zeros(n * n, complex)
    xo, hist_o, conv_o, _ = gmres_oracle(xo, lambda v: O.fastconvolution(Mo, v), rhs_o, Pl_ldiv=Po.solve, maxiter=80)
    xg = np.zeros(n * n, complex)
    xg, hg = ls.gmres_(xg, Mg, rhs_o, Pl=Pg, log=True, maxiter=80)
    assert hg.iters == len(hist_o) and hg.isconverged == conv_o
    m = min(50, len(hist_o))
    assert np.max(np.abs(hg["resnorm"][:m] - hist_o[:m]) / hist_o[:m]) < 1e-8
    assert _rel(xg, xo) < 1e-7


@pytest.mark.parametrize("orth", ["ClassicalGramSchmidt", "DGKS"])
def test_gmres_other_orthogonalisations(orth):
    """orth_meth = ClassicalGramSchmidt / DGKS (IterativeSolvers.jl options; BLAS-2 style sweeps on the GPU)."""
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres as gmres_oracle
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 64
    x, h, k, Mo = O.pow2_problem_2d(n, ppw=6.0, nu=lambda X, Y: 4.0 * O.nu_gaussian_2d(X, Y))
    N = n * n
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega, quadRule="Greengard_Vico")
    rng = np.random.default_rng(21)
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    xo = np.zeros(N, complex)
    xo, hist_o, conv_o, mv_o = gmres_oracle(xo, lambda v: O.fastconvolution(Mo, v), rhs, maxiter=45, reltol=1e-10, orth_meth=orth)
    xg = np.zeros(N, complex)
    xg, hg = ls.gmres_(xg, Mg, rhs, maxiter=45, reltol=1e-10, log=True, orth_meth=orth)
    assert hg.iters == len(hist_o) and hg.mvps == mv_o
    assert np.max(np.abs(hg["resnorm"] - hist_o) / hist_o) < 1e-8
    assert _rel(xg, xo) < 1e-8
    # and a plain multi-column step against numpy (k = 11 crosses the 8-column pass boundary)
    K = ls.KrylovWorkspace(N)
    from fast_solver_lippmann_schwinger_b200._lib import check, lib
    check(lib().ls_krylov_set_orth(K.handle, 1 if orth == "ClassicalGramSchmidt" else 2))
    Q, _ = np.linalg.qr(rng.standard_normal((N, 11)) + 1j * rng.standard_normal((N, 11)))
    w = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    dV = ls.DeviceBuffer.from_host(np.asfortranarray(Q).reshape(-1, order="F"))
    dw = ls.DeviceBuffer.from_host(w)
    hcol = K.mgs_step(dV, N, 11, dw)
    href = Q.conj().T @ w
    wr = w - Q @ href
    assert np.abs(hcol[:11] - href).max() < 1e-12 * np.abs(href).max()
    assert abs(hcol[11] - np.linalg.norm(wr)) < 1e-12 * np.linalg.norm(wr)
    assert _rel(dw.to_host(), wr / np.linalg.norm(wr)) < 1e-12


def test_gmres_options_and_error_paths():
    """initially_zero, abstol, restart bounds, a failing Msp callback, two operators alive at once."""
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres as gmres_oracle
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 64
    x, h, k, Mo = O.pow2_problem_2d(n)
    N = n * n
    A1 = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, n, k, quadRule="Greengard_Vico")
    A2 = ls.FastM(Mo.GFFT, 2.0 * Mo.nu, Mo.ne, Mo.me, n, n, k, quadRule="Greengard_Vico")     # a second handle, own stream
    rng = np.random.default_rng(17)
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    # initially_zero skips the first mul! (mv_products starts at 1 upstream)
    xo, ho, co, mvo = gmres_oracle(np.zeros(N, complex), lambda v: O.fastconvolution(Mo, v), rhs, maxiter=15, initially_zero=True)
    xg, hg = ls.gmres_(np.zeros(N, complex), A1, rhs, maxiter=15, log=True, initially_zero=True)
    assert hg.mvps == mvo and hg.iters == len(ho)
    assert np.max(np.abs(hg["resnorm"] - ho) / ho) < 1e-8
    # abstol stops early; restart = 64 is the largest basis served
    xg, hg = ls.gmres_(np.zeros(N, complex), A1, rhs, abstol=1e30, log=True)
    assert hg.iters == 0 and hg.isconverged
    xg, hg64 = ls.gmres_(np.zeros(N, complex), A1, rhs, restart=64, maxiter=70, reltol=1e-12, log=True)
    xo, ho64, _, _ = gmres_oracle(np.zeros(N, complex), lambda v: O.fastconvolution(Mo, v), rhs, restart=64, maxiter=70, reltol=1e-12)
    assert hg64.iters == len(ho64)
    sig = ho64 > 1e-8 * ho64[0]          # below ~1e-8 of the start the estimate is dominated by rounding in both codes
    assert np.max(np.abs(hg64["resnorm"][sig] - ho64[sig]) / ho64[sig]) < 1e-8
    with pytest.raises(ls.LSUnsupported):
        ls.gmres_(np.zeros(N, complex), A1, rhs, restart=65, maxiter=3)
    # interleaved applies on two handles give independent, correct results
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    y1, y2 = A1 * b, A2 * b
    assert _rel(y1, O.fastconvolution(Mo, b)) < 1e-12
    assert _rel((y2 - b), 2.0 * (y1 - b)) < 1e-12
    # a preconditioner whose host solve fails must surface as an error, not a wrong answer
    import scipy.sparse as sp
    P = ls.SparsifyingPreconditioner(sp.identity(N, dtype=complex, format="csc"), sp.identity(N, dtype=complex, format="csc"))

    class Boom:
        def solve(self, v):
            raise RuntimeError("factorisation lost")
    P.MspInv = Boom()
    with pytest.raises(ls.LSCudaError) as ei:
        ls.gmres_(np.zeros(N, complex), A1, rhs, Pl=P, maxiter=3)
    assert ei.value.code == -6
    # wrong-size vectors are DimensionMismatch-like errors on the host side
    with pytest.raises(ValueError):
        ls.gmres_(np.zeros(N - 1, complex), A1, rhs, maxiter=3)
