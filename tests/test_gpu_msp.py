"""GPU parity of the device-resident Msp^-1 (ls_msp_factor / ls_msp_solve / ls_gmres_msp) against the oracle's
`MspInv = lu(Msp)` (preconditioner.jl:35; SuperLU here) and of the fully device-resident preconditioned GMRES.

Tolerances: a direct solve agrees with SuperLU to 1e-10 relative (both are backward-stable LU factorisations of the
same matrix with different pivot orders); preconditioned GMRES residual histories agree with the oracle to 1e-8
relative over the first 50 iterations (BASELINE.json north_star).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def stencil9(n, m, seed=0, shift=6.0):
    """Random complex 9-point matrix on the n x m grid (x fastest), position-dependent values, safely non-singular."""
    rng = np.random.default_rng(seed)
    I, J = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    I, J = I.reshape(-1, order="F"), J.reshape(-1, order="F")
    row = I + n * J
    rows, cols, vals = [], [], []
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            ok = (I + dx >= 0) & (I + dx < n) & (J + dy >= 0) & (J + dy < m)
            v = rng.standard_normal(ok.sum()) + 1j * rng.standard_normal(ok.sum())
            if dx == 0 and dy == 0:
                v = v + shift
            rows.append(row[ok]); cols.append(row[ok] + dx + n * dy); vals.append(v)
    N = n * m
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))
    A.sort_indices()
    return A


@pytest.mark.parametrize("n,m", [(1, 1), (3, 2), (5, 5), (9, 4), (17, 33), (40, 25), (64, 64), (201, 201), (130, 257)])
def test_msp_solve_matches_superlu(n, m):
    import fast_solver_lippmann_schwinger_b200 as ls
    A = stencil9(n, m, seed=n * 1000 + m)
    N = n * m
    rng = np.random.default_rng(3)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    F = ls.GPUMspFactorization(A, n, m)
    x = F.solve(b)
    x_ref = spla.splu(A).solve(b)
    assert _rel(x, x_ref) < 1e-10
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-11
    # device pointers, in place, and a second right-hand side through the same factorisation
    db = ls.DeviceBuffer.from_host(2.0 * b)
    F.solve(db)
    F.sync()
    assert _rel(db.to_host(), 2.0 * x_ref) < 1e-10
    assert F.factor_bytes > 0 and F.depth >= 0
    F.destroy()


# solver 1 = the uniform-batch kernels of the first version; solver 2 = tightly packed ragged batches (msp_gemv.cuh) with
# the gather kernel kept (fuse 0) or fused into the sweeps, launch geometry by the fixed rule (tune 0) or timed (tune 1)
MSP_CONFIGS = [("1", "0", "0"), ("2", "0", "0"), ("2", "1", "0"), ("2", "0", "1"), ("2", "1", "1")]


@pytest.mark.parametrize("solver,fuse,tune", MSP_CONFIGS)
def test_msp_solver_variants_agree(solver, fuse, tune, monkeypatch):
    """Every solver configuration against SuperLU on ragged dissections (odd and non-square grids, grids smaller than a
    leaf), in place, with the CUDA-graph replay and without it."""
    import fast_solver_lippmann_schwinger_b200 as ls
    monkeypatch.setenv("LS_MSP_SOLVER", solver)
    monkeypatch.setenv("LS_MSP_FUSE", fuse)
    monkeypatch.setenv("LS_MSP_TUNE", tune)
    for (n, m), graph in (((1, 1), "1"), ((3, 2), "1"), ((9, 4), "0"), ((17, 33), "1"), ((40, 25), "0"), ((130, 257), "1"), ((201, 201), "1")):
        monkeypatch.setenv("LS_MSP_GRAPH", graph)
        A = stencil9(n, m, seed=7 * n + m)
        rng = np.random.default_rng(n + m)
        b = rng.standard_normal(n * m) + 1j * rng.standard_normal(n * m)
        x_ref = spla.splu(A).solve(b)
        F = ls.GPUMspFactorization(A, n, m)
        plan = F.plan()
        assert plan.startswith("solver %s " % solver), plan.splitlines()[0]
        if solver == "2":
            assert ("fuse %s tune %s" % (fuse, tune)) in plan.splitlines()[0]
        assert _rel(F.solve(b), x_ref) < 1e-10, (n, m, plan.splitlines()[0])
        db, dx = ls.DeviceBuffer.from_host(b), ls.DeviceBuffer(b.nbytes)
        for _ in range(3):                                  # the first call captures the graph, the next ones replay it
            F.solve(db, dx)
        F.sync()
        assert _rel(dx.to_host(), x_ref) < 1e-10
        F.solve(db)                                         # in place
        F.sync()
        assert _rel(db.to_host(), x_ref) < 1e-10
        db.free(); dx.free()
        F.destroy()


def test_msp_tight_packing_drops_the_padding():
    """Solver 2 stores the blocks with the nodes' actual sizes: fewer bytes than the identity-padded uniform batches."""
    import os
    import fast_solver_lippmann_schwinger_b200 as ls
    n, m = 130, 257
    A = stencil9(n, m, seed=5)
    sizes = {}
    for solver in ("1", "2"):
        os.environ["LS_MSP_SOLVER"] = solver
        try:
            F = ls.GPUMspFactorization(A, n, m)
        finally:
            del os.environ["LS_MSP_SOLVER"]
        sizes[solver] = F.factor_bytes
        F.destroy()
    assert 0 < sizes["2"] < sizes["1"]


@pytest.mark.parametrize("leaf", ["3", "8"])
def test_msp_leaf_size_does_not_change_the_answer(leaf, monkeypatch):
    import fast_solver_lippmann_schwinger_b200 as ls
    monkeypatch.setenv("LS_MSP_LEAF", leaf)
    n, m = 37, 52
    A = stencil9(n, m, seed=11)
    rng = np.random.default_rng(4)
    b = rng.standard_normal(n * m) + 1j * rng.standard_normal(n * m)
    F = ls.GPUMspFactorization(A, n, m)
    assert _rel(F.solve(b), spla.splu(A).solve(b)) < 1e-10


def test_msp_rejects_what_it_does_not_serve():
    import fast_solver_lippmann_schwinger_b200 as ls
    n, m = 12, 9
    A = stencil9(n, m, seed=2).tolil()
    A[0, 5] = 1.0                                 # couples (0,0) with (5,0): outside the 9-point pattern
    with pytest.raises(ls.LSUnsupported):
        ls.GPUMspFactorization(A.tocsc(), n, m)
    with pytest.raises(ValueError):
        ls.GPUMspFactorization(stencil9(n, m), n, m + 1)
    # a singular pivot block must surface as an error, not as NaNs
    Z = sp.csc_matrix((n * m, n * m), dtype=complex)
    with pytest.raises(ls.LSCudaError):
        ls.GPUMspFactorization(Z, n, m)


@pytest.mark.parametrize("n", [64, 201])
def test_device_preconditioned_gmres_matches_oracle(n):
    """examples/example.jl as shipped (n = 201) and a power-of-two grid: gmres!(u, fastconv, rhs, Pl=precond) with
    As*b and Msp^-1 both on the GPU against the oracle's SuperLU-preconditioned history."""
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres as gmres_oracle
    import fast_solver_lippmann_schwinger_b200 as ls
    h = 1.0 / n if n % 2 == 0 else 0.005
    x = (-0.5 + h * np.arange(n)) if n % 2 == 0 else (-0.5 + h * np.arange(n))
    k = 1.0 / h
    Mo = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    X, Y = O.grid2d(x, x)
    D0 = O.referenceValsTrapRule()[1][0]
    cache = O.entriesSparseA(k, X, Y, D0, n, n, strict=False)
    As = O.buildSparseA(k, X, Y, D0, n, n, strict=False, _cache=cache)
    AG = O.buildSparseAG(k, X, Y, D0, n, n, strict=False, _cache=cache)
    Msp = (As + k ** 2 * (AG @ sp.diags(Mo.nu))).tocsc()
    N = n * n
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, n, k, quadRule="Greengard_Vico")
    u_inc = np.exp(1j * k * X)
    rhs = -k ** 2 * O.FFTconvolution(Mo, Mo.nu * u_inc)            # examples/example.jl:76-77
    Po = O.SparsifyingPreconditioner(Msp, As)
    Pg = ls.SparsifyingPreconditioner(Msp, As, solverType="GPU", grid=(n, n))
    # M \ b on the device against the oracle's
    rng = np.random.default_rng(8)
    v = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    assert _rel(Pg.solve(v), Po.solve(v)) < 1e-9
    xo = np.zeros(N, complex)
    xo, hist_o, conv_o, mv_o = gmres_oracle(xo, lambda w: O.fastconvolution(Mo, w), rhs, Pl_ldiv=Po.solve, maxiter=60)
    xg = np.zeros(N, complex)
    xg, hg = ls.gmres_(xg, Mg, rhs, Pl=Pg, log=True, maxiter=60)
    assert hg.iters == len(hist_o) and hg.isconverged == conv_o and hg.mvps == mv_o
    mm = min(50, len(hist_o))
    assert np.max(np.abs(hg["resnorm"][:mm] - hist_o[:mm]) / hist_o[:mm]) < 1e-8
    assert _rel(xg, xo) < 1e-8
    assert hg.msp_host_seconds == 0.0                              # nothing crossed PCIe inside the loop
    # the host-callback route gives the same history and reports its host time
    Ph = ls.SparsifyingPreconditioner(Msp, As)
    xh, hh = ls.gmres_(np.zeros(N, complex), Mg, rhs, Pl=Ph, log=True, maxiter=60)
    assert hh.iters == hg.iters and np.max(np.abs(hh["resnorm"] - hg["resnorm"]) / hg["resnorm"]) < 1e-8
    assert hh.msp_host_seconds > 0.0


def test_gmres_maxiter_semantics_follow_upstream():
    """maxiter = 0 runs no iteration; maxiter inside a cycle returns the x of the last restart (gmres.jl iterate)."""
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres as gmres_oracle
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 64
    x, h, k, Mo = O.pow2_problem_2d(n)
    N = n * n
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, n, k, quadRule="Greengard_Vico")
    rng = np.random.default_rng(2)
    rhs = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    x0 = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    xg, hg = ls.gmres_(x0.copy(), Mg, rhs, maxiter=0, log=True)
    assert hg.iters == 0 and np.array_equal(xg, x0)
    xo, ho, co, mvo = gmres_oracle(x0.copy(), lambda v: O.fastconvolution(Mo, v), rhs, restart=5, maxiter=7, reltol=1e-14)
    xg, hg = ls.gmres_(x0.copy(), Mg, rhs, restart=5, maxiter=7, reltol=1e-14, log=True)
    assert hg.iters == 7 and hg.mvps == mvo
    assert _rel(xg, xo) < 1e-10                                    # both hold the iterate of the restart at 5
    x5, _ = ls.gmres_(x0.copy(), Mg, rhs, restart=5, maxiter=5, reltol=1e-14, log=True)
    assert _rel(xg, x5) < 1e-12


def test_sparsifier2d_sampled_and_reduced_on_the_device():
    """SURVEY 8(f) row 2 in 2-D: buildSparseAConv / buildSparseAGConv (SparsifyingMatrix2D.jl:104-201, 278-350, 441-532, 888-966)
    with the 49 unit-vector FFTconvolution applies, the rows and their far-field Gram matrices kept on the GPU
    (ls_sample_rows / ls_gram / ls_gather_rows) against the oracle's QR + SVD of the host-sampled rows, up to the per-row
    phase (Q5); then the whole preconditioned solve on the device with these matrices."""
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200 import sparsifier as S
    n = 64
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    k = 2 * np.pi / (8.3 * h)
    Mo = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    X, Y = O.grid2d(x, x)
    cache = O.entriesSparseAConv(k, X, Y, Mo, n, n, strict=False)
    As_o = O.buildSparseAConv(k, X, Y, Mo, n, n, strict=False, _cache=cache)
    AG_o = O.buildSparseAGConv(k, X, Y, Mo, n, n, strict=False, _cache=cache)
    Msp_o = (As_o + k ** 2 * (AG_o @ sp.diags(Mo.nu))).tocsc()
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, n, k, quadRule="Greengard_Vico")
    l0 = Mg.launch_count()
    As, Msp = S.sparsifying_matrices_2d(k, X, Y, Mg, n, n, Mo.nu, strict=False)
    assert Mg.launch_count() - l0 == 3 * (9 + 4 * 6 + 4 * 4)          # one sampling pass: 49 applies of three kernels
    A, B = As.tocsr(), As_o.tocsr()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    for r in range(0, A.shape[0], 13):
        a, b = A.data[A.indptr[r]:A.indptr[r + 1]], B.data[B.indptr[r]:B.indptr[r + 1]]
        ph = np.vdot(b, a) / abs(np.vdot(b, a))
        assert np.abs(a - ph * b).max() <= 1e-7 * np.abs(b).max(), r
    v = np.random.default_rng(5).standard_normal(n * n) + 0j
    w_o = spla.splu(Msp_o).solve(As_o @ v)
    assert _rel(spla.splu(Msp).solve(As @ v), w_o) < 1e-6
    P = ls.SparsifyingPreconditioner(Msp, As, solverType="GPU", grid=(n, n))
    assert _rel(P.solve(v), w_o) < 1e-6
    u_inc = np.exp(1j * k * X)
    rhs = -(Mg * u_inc - u_inc)
    u, hist = ls.gmres_(np.zeros(n * n, complex), Mg, rhs, Pl=P, log=True, reltol=1e-8)
    assert hist.isconverged
    assert np.linalg.norm((Mg * u) - rhs) / np.linalg.norm(rhs) < 1e-6
