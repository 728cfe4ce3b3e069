"""Host logic of fast_solver_lippmann_schwinger_b200.sparsifier against the oracle, with a CPU
apply injected in place of the GPU operator (the package itself never imports the oracle)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import ls_oracle as O


def _phase_aligned_equal(A, B, tol=1e-9):
    """Rows of A and B agree up to one unit phase per row (singular vectors are phase-ambiguous)."""
    A, B = A.tocsr(), B.tocsr()
    assert A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    for r in range(A.shape[0]):
        a, b = A.data[A.indptr[r]:A.indptr[r + 1]], B.data[B.indptr[r]:B.indptr[r + 1]]
        ph = np.vdot(b, a)
        ph /= abs(ph)
        assert np.abs(a - ph * b).max() <= tol * np.abs(b).max(), r
    return True


def test_sparsifier_3d_matches_oracle():
    from fast_solver_lippmann_schwinger_b200 import sparsifier as S
    n, l = 8, 10
    (x, z), h, k, Mo, As_o, Msp_o, Po = O.example_problem_3d(n, l)
    X, Y, Z = O.grid3d(x, x, z)
    calls = []

    def apply(M, e):
        calls.append(1)
        return O.FFTconvolution3D(Mo, e)

    As, Msp = S.sparsifying_matrices_3d(k, X, Y, Z, object(), n, n, l, Mo.nu, apply=apply)
    assert len(calls) == 27 + 6 * 18 + 12 * 12 + 8 * 8            # one sampling pass: 343 unit-vector applies
    assert _phase_aligned_equal(As, As_o) and _phase_aligned_equal(Msp, Msp_o)
    # the preconditioner does not see the phases
    v = np.random.default_rng(1).standard_normal(n * n * l) + 1j * np.random.default_rng(2).standard_normal(n * n * l)
    w = spla.splu(Msp).solve(As @ v)
    assert np.linalg.norm(w - Po.solve(v)) <= 1e-9 * np.linalg.norm(w)
    # upstream-named entry points
    Ind, Val = S.entriesSparseA3D(k, X, Y, Z, object(), n, n, l, apply=apply)
    assert [len(i) for i in Ind] == [27] + [18] * 6 + [12] * 12 + [8] * 8
    A2 = S.buildSparseA3DConv(k, X, Y, Z, object(), n, n, l, apply=apply)
    assert _phase_aligned_equal(A2, As_o)
    with pytest.raises(IndexError):
        S.buildSparseA3DConv(k, X[:8], Y[:8], Z[:8], object(), 2, 2, 2, apply=lambda M, e: e)   # no interior point


@pytest.mark.parametrize("n,strict", [(15, True), (16, False)])
def test_sparsifier_2d_conv_matches_oracle(n, strict):
    from fast_solver_lippmann_schwinger_b200 import sparsifier as S
    h = 1.0 / (n - 1) if n % 2 else 1.0 / n
    x = -0.5 + h * np.arange(n)
    k = 2 * np.pi / (8.3 * h)          # 8 points per wavelength on n = 16 hits Gtruncated2D's s == k singularity (Q7)
    Mo = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    X, Y = O.grid2d(x, x)
    cache = O.entriesSparseAConv(k, X, Y, Mo, n, n, strict=strict)
    As_o = O.buildSparseAConv(k, X, Y, Mo, n, n, strict=strict, _cache=cache)
    AG_o = O.buildSparseAGConv(k, X, Y, Mo, n, n, strict=strict, _cache=cache)

    def apply(M, e):
        return O.FFTconvolution(Mo, e)

    pc = S.entriesSparseAConv(k, X, Y, object(), n, n, apply=apply, strict=strict)
    As = S.buildSparseAConv(k, X, Y, object(), n, n, apply=apply, strict=strict, _cache=pc)
    AG = S.buildSparseAGConv(k, X, Y, object(), n, n, apply=apply, strict=strict, _cache=pc)
    assert As.nnz == 9 * (n - 2) ** 2 + 6 * 4 * (n - 2) + 4 * 4
    assert _phase_aligned_equal(As, As_o)
    Msp = (As + k ** 2 * (AG @ sp.diags(Mo.nu))).tocsc()
    Msp_o = (As_o + k ** 2 * (AG_o @ sp.diags(Mo.nu))).tocsc()
    v = np.random.default_rng(4).standard_normal(n * n) + 0j
    w, w_o = spla.splu(Msp).solve(As @ v), spla.splu(Msp_o).solve(As_o @ v)
    assert np.linalg.norm(w - w_o) <= 1e-8 * np.linalg.norm(w_o)
    if strict:
        with pytest.raises(AssertionError):
            S.entriesSparseAConv(k, X, Y, object(), 16, 16, apply=apply)          # upstream asserts odd sizes
