"""GPU tests of the 3-D sparsifier path (run last: file name sorts after the other GPU tests).

* the row-slab storage of the sharded SpMV on one GPU (window columns, no exchange);
* examples/example3D.jl:56-78 end to end: 27-point As and Msp from the oracle's restatement of
  SparsifyingMatrix3D.jl, As on the GPU, the Msp solve on the host, GMRES histories against the oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("classes", [True, False])
def test_row_slab_sparse_matrix_on_one_gpu(classes):
    """ls_spm_create_dist on an unsharded 3-D operator: the windowed row block [halo | rows | halo] without any
    exchange - stencil-class and CSR storage - against scipy."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from util_sparse import stencil27
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200 import dist as lsd
    n = l = 64
    h = 1.0 / n
    k = 2 * np.pi / (10 * h)
    nu = np.zeros(n * n * l)
    M = lsd.FastM3DSharded(nu, n, n, l, k, 1.8 * n * h, 4.0 * n * h, 0, 1, None)
    A = stencil27(n, n, l, seed=11, classes=classes)
    As = lsd.GPUSparseMatrixCSCSharded(A, M)
    assert As.halo == n * n + n + 1
    assert As.format == ("stencil" if classes else "csr")
    if classes:
        assert As.nclasses == 27
    rng = np.random.default_rng(9)
    x = rng.standard_normal(n * n * l) + 1j * rng.standard_normal(n * n * l)
    ref = A @ x
    y = As * x
    assert _rel(y, ref) <= 1e-14
    y2 = As.mv(x, y=y.copy(), alpha=0.5 - 1j, beta=2.0)
    assert _rel(y2, (0.5 - 1j) * ref + 2.0 * y) <= 1e-14
    dx, dy = ls.DeviceBuffer.from_host(x), ls.DeviceBuffer(x.nbytes)
    As.mv(dx, dy)
    As.sync()                                     # device-pointer calls are asynchronous on the handle's stream
    assert _rel(dy.to_host(), ref) <= 1e-14
    with pytest.raises(ls.LSCudaError):
        lsd.GPUSparseMatrixCSCSharded(A, M, halo=n * n * l + 1)     # halo larger than the slab


def test_example3d_preconditioned_gmres_matches_oracle():
    """gmres!(u, fastconv, rhs, precond) of examples/example3D.jl:56-78 on a 20 x 20 x 36 grid (the script ships 48^3;
    the host LU of Msp is what limits the test size): operator on the general-size GPU path, As (27 stencil classes)
    on the GPU, Msp^-1 through the host callback; residual history against the oracle."""
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres as gmres_oracle
    import fast_solver_lippmann_schwinger_b200 as ls
    n, l = 20, 36
    (x, z), h, k, Mo, As, Msp, Po = O.example_problem_3d(n, l)
    N = n * n * l
    assert As.nnz == (n - 2) ** 2 * (l - 2) * 27 + (2 * (n - 2) ** 2 + 4 * (n - 2) * (l - 2)) * 18 \
        + (8 * (n - 2) + 4 * (l - 2)) * 12 + 64
    X, Y, Z = O.grid3d(x, x, z)
    u_inc = np.exp(1j * k * X)
    rhs = -(Mo * u_inc - u_inc)                                   # example3D.jl:71-72
    uo, hist_o, conv_o, mv_o = gmres_oracle(np.zeros(N, complex), lambda v: Mo * v, rhs, Pl_ldiv=Po.solve)
    assert conv_o

    Mg = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k)
    Pg = ls.SparsifyingPreconditioner(Msp, As)
    assert Pg.As.format == "stencil" and Pg.As.nclasses == 27
    v = np.random.default_rng(3).standard_normal(N) + 0j
    assert _rel(Pg.As * v, As @ v) <= 1e-14
    ug, hg = ls.gmres_(np.zeros(N, complex), Mg, rhs, Pl=Pg, log=True)
    assert hg.isconverged and hg.iters == len(hist_o)
    assert np.max(np.abs(hg["resnorm"] - hist_o) / hist_o) < 1e-8
    assert _rel(ug, uo) < 1e-7


def test_sparsifier_built_from_gpu_applies():
    """As / Msp of examples/example3D.jl:56-61 sampled through FFTconvolution on the GPU operator
    (343 unit-vector applies), against the oracle's matrices up to the per-row phase (Q5)."""
    from oracle import ls_oracle as O
    import scipy.sparse.linalg as spla
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200 import sparsifier as S
    n, l = 20, 36
    (x, z), h, k, Mo, As_o, Msp_o, Po = O.example_problem_3d(n, l)
    X, Y, Z = O.grid3d(x, x, z)
    Mg = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k)
    l0 = Mg.launch_count()
    As, Msp = S.sparsifying_matrices_3d(k, X, Y, Z, Mg, n, n, l, Mo.nu)
    assert Mg.launch_count() > l0                                  # the rows came from the device
    assert As.nnz == As_o.nnz and Msp.nnz == Msp_o.nnz
    A, B = As.tocsr(), As_o.tocsr()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    for r in range(0, A.shape[0], 97):                             # per-row unit phase
        a, b = A.data[A.indptr[r]:A.indptr[r + 1]], B.data[B.indptr[r]:B.indptr[r + 1]]
        ph = np.vdot(b, a) / abs(np.vdot(b, a))
        assert np.abs(a - ph * b).max() <= 1e-7 * np.abs(b).max()
    v = np.random.default_rng(8).standard_normal(n * n * l) + 0j
    w = spla.splu(Msp).solve(As @ v)
    assert _rel(w, Po.solve(v)) <= 1e-7
