"""Property tests (hypothesis) of the host-side logic that feeds the GPU path: sparsifier assembly, slab bookkeeping, the
row-slab blocks of the sharded SpMV.  No GPU."""
import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings, strategies as st

from fast_solver_lippmann_schwinger_b200 import dist as lsd
from fast_solver_lippmann_schwinger_b200 import sparsifier as S


def _coo_reference(N, row_sets, Indices, Values):
    """sparse(row, col, val) of createIndices' triplets (Functions.jl:7-29), the way upstream assembles."""
    rows, cols, vals = [], [], []
    for rset, ind, val in zip(row_sets, Indices, Values):
        R, Cc, V = S.createIndices(rset, ind, val)
        rows.append(R); cols.append(Cc); vals.append(V)
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows) - 1, np.concatenate(cols) - 1)), shape=(N, N)).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    return A


@settings(max_examples=40, deadline=None)
@given(st.integers(3, 9), st.integers(3, 9), st.integers(0, 2 ** 31 - 1))
def test_assemble_equals_triplet_assembly_2d(n, m, seed):
    """Row-by-row assembly == sparse() of the createIndices triplets, for the nine 2-D boundary classes of an n x m grid
    with random stencil values (also in a scrambled stencil order, which upstream's edge classes have)."""
    rng = np.random.default_rng(seed)
    rows = S._rows2(n, m)
    Ind, Val = [], []
    for centre, relA, relG in S._classes_2d(n, m, False):
        rel = np.asarray(relA, dtype=np.int64)
        perm = rng.permutation(rel.size)
        Ind.append(rel[perm])
        Val.append(rng.standard_normal(rel.size) + 1j * rng.standard_normal(rel.size))
    A = S._assemble(n * m, rows, Ind, Val)
    B = _coo_reference(n * m, rows, Ind, Val)
    assert A.format == "csc" and A.has_sorted_indices
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)
    # every row of the grid has exactly its in-grid neighbours
    cnt = np.diff(A.tocsr().indptr).reshape((n, m), order="F")
    assert cnt[1:-1, 1:-1].min() == 9 and cnt[0, 0] == 4 and cnt[0, 1:-1].max() == 6


@settings(max_examples=25, deadline=None)
@given(st.integers(3, 6), st.integers(3, 6), st.integers(3, 6), st.integers(0, 2 ** 31 - 1))
def test_assemble_equals_triplet_assembly_3d(n, m, l, seed):
    rng = np.random.default_rng(seed)
    rows = S._rows3(n, m, l)
    Ind = [S._rel3(cls, n, m) for cls in S._CLASSES_3D]
    Val = [rng.standard_normal(i.size) + 1j * rng.standard_normal(i.size) for i in Ind]
    A = S._assemble(n * m * l, rows, Ind, Val)
    B = _coo_reference(n * m * l, rows, Ind, Val)
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)
    assert A.nnz == (3 * n - 2) * (3 * m - 2) * (3 * l - 2)          # product of the 1-D tridiagonal counts


@settings(max_examples=40, deadline=None)
@given(st.integers(3, 8), st.integers(3, 8), st.integers(0, 2 ** 31 - 1), st.floats(0.5, 50.0))
def test_system_matrix_one_pass_equals_sparse_algebra(n, m, seed, k):
    """Mapproxsp = As + k^2 AG diag(nu) (examples/example.jl:67) from the shared pattern == scipy's sparse algebra."""
    rng = np.random.default_rng(seed)
    rows = S._rows2(n, m)
    Ind = [np.asarray(relA, dtype=np.int64) for c, relA, relG in S._classes_2d(n, m, False)]
    mk = lambda: [rng.standard_normal(i.size) + 1j * rng.standard_normal(i.size) for i in Ind]
    As, AG = S._assemble(n * m, rows, Ind, mk()), S._assemble(n * m, rows, Ind, mk())
    nu = rng.standard_normal(n * m)
    M1 = S._system_matrix(As, AG, k, nu)
    M2 = (As + k ** 2 * (AG @ sp.diags(nu))).tocsc()
    assert abs(M1 - M2).max() <= 1e-12 * max(1.0, abs(M2).max())
    # a different pattern takes the general route
    AG2 = AG.copy().tolil(); AG2[0, n * m - 1] = 1.0
    M3 = S._system_matrix(As, AG2.tocsc(), k, nu)
    assert abs(M3 - (As + k ** 2 * (AG2.tocsc() @ sp.diags(nu)))).max() <= 1e-12 * max(1.0, abs(M2).max())


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 6), st.integers(1, 6), st.sampled_from([1, 2, 4, 8]), st.integers(1, 4))
def test_slabs_partition_the_grid_vector(n, m, P, planes_per_rank):
    l = P * planes_per_rank
    pieces = [lsd.vector_range(n, m, l, r, P) for r in range(P)]
    assert pieces[0][0] == 0 and pieces[-1][1] == n * m * l
    assert all(pieces[r][1] == pieces[r + 1][0] for r in range(P - 1))
    assert all(b - a == n * m * planes_per_rank for a, b in pieces)
    v = np.arange(n * m * l, dtype=complex)
    assert np.array_equal(np.concatenate([lsd.scatter_vector(v, n, m, l, r, P) for r in range(P)]), v)


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 5), st.integers(2, 5), st.sampled_from([2, 4]), st.integers(2, 3), st.integers(0, 2 ** 31 - 1))
def test_row_slab_blocks_reproduce_the_product(n, m, P, planes_per_rank, seed):
    """Every rank's windowed row block times [halo | slab | halo] == its slab of A x, for a random 27-point matrix."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from util_sparse import stencil27
    l = P * planes_per_rank
    A = stencil27(n, m, l, seed=seed % 1000)
    H = lsd.matrix_halo(A)
    N = n * m * l
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    y = A @ x
    for r in range(P):
        a, b = lsd.vector_range(n, m, l, r, P)
        if H > b - a:
            continue                                   # one plane per rank is thinner than the stencil reach: rejected elsewhere
        blk = lsd.local_block_csc(A, a, b, H)
        xext = np.zeros(b - a + 2 * H, complex)
        lo, hi = max(a - H, 0), min(b + H, N)
        xext[lo - (a - H): lo - (a - H) + hi - lo] = x[lo:hi]
        assert np.allclose(blk @ xext, y[a:b], rtol=0, atol=1e-12)
