"""Sparsifying matrices built from operator applies.

The reference builds the preconditioner's stencil matrix ``As`` by sampling rows of the discrete
Green's operator with ``FFTconvolution`` applied to unit vectors and taking, per boundary class, the
last left singular vector of the far-field block (2-D: sampleGConv FastConvolution.jl:278-306,
entriesSparseAConv / entriesSparseGConv / buildSparseAGConv / buildSparseAConv
SparsifyingMatrix2D.jl:104-201, 278-350, 441-532, 888-966; 3-D: sampleG3D FastConvolution3D.jl:136-160,
entriesSparseG3D / entriesSparseA3D / buildSparseA3DConv / buildSparseAG3DConv
SparsifyingMatrix3D.jl:963-1135, 1136-1408, 1410-1653, 1659-1917).  Here the applies run on the GPU
operator (``FastM`` / ``FastM3D``); the 9 / 27 small SVDs stay on the host.

Same names, argument order and return values as upstream.  The stencil vectors are singular vectors and
therefore defined up to a unit phase: ``Msp^-1 As`` does not depend on it.

Every function takes ``apply=`` (default: ``FFTconvolution`` of this package, i.e. the GPU path); tests
inject a CPU apply to check the bookkeeping without a device.  Host memory: a class samples
``len(stencil) x N`` complex numbers (27 N at most: 0.9 GB at 128^3, 7.2 GB at 256^3).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .operators import FFTconvolution as _gpu_fftconvolution

__all__ = ["createIndices", "sampleGConv", "sampleG3D", "entriesSparseAConv", "entriesSparseGConv",
           "buildSparseAConv", "buildSparseAGConv", "sparsifying_matrices_2d", "entriesSparseA3D", "entriesSparseG3D",
           "buildSparseA3DConv", "buildSparseAG3DConv", "sparsifying_matrices_3d"]


def createIndices(row, col, val):
    """Functions.jl:7-29: every row gets the same (relative column, value) list.  1-based indices."""
    row = np.atleast_1d(np.asarray(row, dtype=np.int64))
    col = np.asarray(col, dtype=np.int64).reshape(-1)
    val = np.asarray(val, dtype=np.complex128).reshape(-1)
    if col.shape != val.shape:
        raise AssertionError("length(col) == length(val)")
    Row = np.repeat(row, col.size)
    return Row, np.tile(col, row.size) + Row, np.tile(val, row.size)


class _HostRows:
    """Rows of G sampled through a host-visible apply (tests inject one): s x N array on the host."""

    def __init__(self, rows, ind):
        self.rows, self.ind = rows, np.asarray(ind, dtype=np.int64)

    def null_vector(self):
        """U[:, end]' of svd(rows[:, far]) - the stencil coefficients (SparsifyingMatrix2D.jl:119-127)."""
        far = np.ones(self.rows.shape[1], dtype=bool)
        far[self.ind - 1] = False
        return _last_left_singular_vector(self.rows[:, far])

    def block(self, perm=None):
        """rows[:, ind] (optionally with rows and columns re-ordered by perm): G restricted to the stencil."""
        B = self.rows[:, self.ind - 1]
        return B if perm is None else B[np.ix_(perm, perm)]

    def free(self):
        self.rows = None


class _DeviceRows:
    """The same on the GPU (ls_sample_rows / ls_gram / ls_gather_rows): the s rows never leave the device; the far-field
    Gram matrix GS GS^H = (full Gram) - (near block)(near block)^H comes back as s^2 numbers and its eigenvector of the
    smallest eigenvalue is the last left singular vector of GS.  (Squaring costs accuracy: with sigma_min / sigma_max
    ~ 1e-2..1e-3 for these blocks the coefficients are good to ~1e-10 instead of 1e-15 - immaterial for a preconditioner
    whose stencils are defined up to a phase anyway)"""

    def __init__(self, fastconv, N, ind, ws):
        import ctypes as C
        from ._lib import DeviceBuffer, check, lib, ptr
        self.ind = np.ascontiguousarray(ind, dtype=np.int64)
        self.s, self.N, self.ws = int(self.ind.size), int(N), ws
        if self.ind.min() < 1 or self.ind.max() > N:
            raise IndexError("stencil index outside the grid (grid too small for a 3-point stencil?)")
        self.V = DeviceBuffer(16 * self.N * self.s)
        check(lib().ls_sample_rows(ws.handle, fastconv.handle, ptr(self.ind), self.s, ptr(self.V), self.N))
        self._B = None

    def _near(self):
        from ._lib import check, lib, ptr
        if self._B is None:
            B = np.empty((self.s, self.s), dtype=np.complex128)
            check(lib().ls_gather_rows(self.ws.handle, ptr(self.V), self.N, self.s, ptr(self.ind), self.s, ptr(B)))
            self._B = B
        return self._B

    def null_vector(self):
        from ._lib import check, lib, ptr
        g = np.empty((self.s, self.s), dtype=np.complex128)          # g[j, i] = sum_c conj(rows[i, c]) rows[j, c]
        check(lib().ls_gram(self.ws.handle, ptr(self.V), self.N, self.s, ptr(g)))
        B = self._near()
        GG = g - B @ B.conj().T                                       # far-field GS GS^H (Hermitian)
        GG = 0.5 * (GG + GG.conj().T)
        w, U = np.linalg.eigh(GG)
        return np.conj(U[:, 0])

    def block(self, perm=None):
        B = self._near()
        return B if perm is None else B[np.ix_(perm, perm)]

    def free(self):
        if self.V is not None:
            self.V.free()
            self.V = None


def _is_gpu_operator(fastconv):
    return hasattr(fastconv, "handle") and hasattr(fastconv, "_apply")


def _sampler(fastconv, N, apply):
    """Row sampler for one sparsifier build: device-resident when the operator is a GPU handle and no apply is injected."""
    if apply is None and _is_gpu_operator(fastconv):
        from .krylov import KrylovWorkspace
        ws = KrylovWorkspace(N)
        return lambda ind: _DeviceRows(fastconv, N, ind, ws)
    return lambda ind: _HostRows(_sample_rows(fastconv, N, ind, apply), ind)


def _sample_rows(fastconv, N, indS, apply):
    """Row i = FFTconvolution(fastconv, e_{indS[i]}); indS is 1-based."""
    apply = _gpu_fftconvolution if apply is None else apply
    indS = np.asarray(indS, dtype=np.int64).reshape(-1)
    if indS.min() < 1 or indS.max() > N:
        raise IndexError("stencil index outside the grid (grid too small for a 3-point stencil?)")
    G = np.empty((indS.size, N), dtype=np.complex128)
    e = np.zeros(N, dtype=np.complex128)
    for i, s0 in enumerate(indS):
        e[s0 - 1] = 1.0
        G[i, :] = apply(fastconv, e)
        e[s0 - 1] = 0.0
    return G


def _last_left_singular_vector(GS):
    """U[:, end]' of svd(GS) for a wide matrix: QR of the tall adjoint, then the SVD of the small R'."""
    R = np.linalg.qr(GS.conj().T, mode="r")
    U, s, Vh = np.linalg.svd(R.conj().T)
    return np.conj(U[:, -1])


def _assemble(N, row_sets, Indices, Values):
    """createIndices (Functions.jl:7-29) over all boundary classes + sparse(row, col, val): every row of a class carries the
    class's (relative column, value) list.  The classes partition the rows, so the matrix is written row by row (CSR) without
    sorting 9 N / 27 N triplets, then converted to the CSC Julia holds."""
    counts = np.zeros(N, dtype=np.int64)
    sets = []
    for rset, ind, val in zip(row_sets, Indices, Values):
        r = np.atleast_1d(np.asarray(rset, dtype=np.int64)).reshape(-1) - 1
        ind = np.asarray(ind, dtype=np.int64).reshape(-1)
        val = np.asarray(val, dtype=np.complex128).reshape(-1)
        if ind.shape != val.shape:
            raise AssertionError("length(col) == length(val)")
        if r.size and (r.min() < 0 or r.max() >= N or r.min() + ind.min() < 0 or r.max() + ind.max() >= N):
            raise IndexError("sparsifier stencil reaches outside the grid")
        if np.any(counts[r] != 0):
            raise AssertionError("boundary classes overlap")
        counts[r] = ind.size
        sets.append((r, ind, val))
    indptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    indices = np.empty(int(indptr[-1]), dtype=np.int64)
    data = np.empty(int(indptr[-1]), dtype=np.complex128)
    for r, ind, val in sets:
        for q in range(ind.size):                # one strided pass per stencil entry: no N x stencil temporaries
            pos = indptr[r] + q
            indices[pos] = r + ind[q]
            data[pos] = val[q]
    A = sp.csr_matrix((data, indices, indptr), shape=(N, N)).tocsc()
    A.sort_indices()
    return A


def _same_pattern(A, B):
    return A.shape == B.shape and A.nnz == B.nnz and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)


def _system_matrix(As, AG, k, nu):
    """Mapproxsp = As + k^2 AG diag(nu) (examples/example.jl:67).  As and AG share their pattern: one pass over the values."""
    nu = np.asarray(nu, dtype=np.float64)
    if As.format == "csc" and AG.format == "csc" and _same_pattern(As, AG):
        col_nu = np.repeat(nu, np.diff(As.indptr))
        return sp.csc_matrix((As.data + (k ** 2) * (AG.data * col_nu), As.indices.copy(), As.indptr.copy()), shape=As.shape)
    return (As + k ** 2 * (AG @ sp.diags(nu))).tocsc()


# ------------------------------------------------------------------------------------------- 3-D
# boundary classes in the order upstream pushes / consumes them: interior, 6 faces, 12 edges, 8 corners;
# per dimension "lo" = first plane (offsets 0,+1), "mid" = interior (-1,0,+1), "hi" = last plane (-1,0)
_CLASSES_3D = (
    ("mid", "mid", "mid"),
    ("lo", "mid", "mid"), ("hi", "mid", "mid"), ("mid", "lo", "mid"), ("mid", "hi", "mid"),
    ("mid", "mid", "lo"), ("mid", "mid", "hi"),
    ("lo", "lo", "mid"), ("hi", "lo", "mid"), ("lo", "hi", "mid"), ("hi", "hi", "mid"),
    ("lo", "mid", "lo"), ("hi", "mid", "lo"), ("lo", "mid", "hi"), ("hi", "mid", "hi"),
    ("mid", "lo", "lo"), ("mid", "hi", "lo"), ("mid", "lo", "hi"), ("mid", "hi", "hi"),
    ("lo", "lo", "lo"), ("hi", "lo", "lo"), ("lo", "hi", "lo"), ("hi", "hi", "lo"),
    ("lo", "lo", "hi"), ("hi", "lo", "hi"), ("lo", "hi", "hi"), ("hi", "hi", "hi"),
)
_OFFSETS = {"lo": (0, 1), "mid": (-1, 0, 1), "hi": (-1, 0)}


def _rel3(cls, n, m):
    cx, cy, cz = cls
    return np.array([dx + n * dy + n * m * dz for dz in _OFFSETS[cz] for dy in _OFFSETS[cy] for dx in _OFFSETS[cx]],
                    dtype=np.int64)


def _centre3(cls, n, m, l):
    def coord(c, nn):
        return 1 if c == "lo" else (nn if c == "hi" else int(round(nn / 2)))    # round half to even, like Julia
    i, j, p = coord(cls[0], n), coord(cls[1], m), coord(cls[2], l)
    return (p - 1) * n * m + (j - 1) * n + i                                   # changeInd3D


def _rows3(n, m, l):
    Ind = np.arange(1, n * m * l + 1, dtype=np.int64).reshape((n, m, l), order="F")
    sel = {"lo": slice(0, 1), "mid": slice(1, -1), "hi": slice(-1, None)}
    return [Ind[sel[cx], sel[cy], sel[cz]].reshape(-1, order="F") for cx, cy, cz in _CLASSES_3D]


def sampleG3D(k, X, Y, Z, indS, fastconv, apply=None):
    """FastConvolution3D.jl:136-160."""
    return _sample_rows(fastconv, len(X), indS, apply)


def _sample_classes_3d(fastconv, n, m, l, apply):
    """One sampling per class, shared by entriesSparseA3D and entriesSparseG3D: (rel, null vector, stencil block).
    With a GPU operator the rows are sampled, reduced (Gram matrix) and dropped class by class on the device."""
    out = []
    sample = _sampler(fastconv, n * m * l, apply)
    for cls in _CLASSES_3D:
        rel = _rel3(cls, n, m)
        ind = _centre3(cls, n, m, l) + rel
        rows = sample(ind)
        out.append((rel, rows.null_vector(), rows.block()))
        rows.free()
    return out


def entriesSparseA3D(k, X, Y, Z, fastconv, n, m, l, apply=None, _samples=None):
    """SparsifyingMatrix3D.jl:1136-1408 -> (Indices, Entries)."""
    samples = _samples if _samples is not None else _sample_classes_3d(fastconv, n, m, l, apply)
    return [rel for rel, v, blk in samples], [v for rel, v, blk in samples]


def entriesSparseG3D(k, X, Y, Z, fastconv, n, m, l, apply=None, _samples=None):
    """SparsifyingMatrix3D.jl:963-1135: G restricted to each class's own stencil."""
    samples = _samples if _samples is not None else _sample_classes_3d(fastconv, n, m, l, apply)
    return [blk for rel, v, blk in samples]


def buildSparseA3DConv(k, X, Y, Z, fastconv, n, m, l, apply=None, _samples=None):
    """SparsifyingMatrix3D.jl:1410-1653 (method "normal"): the 27-point matrix As."""
    Indices, Values = entriesSparseA3D(k, X, Y, Z, fastconv, n, m, l, apply, _samples)
    return _assemble(n * m * l, _rows3(n, m, l), Indices, Values)


def buildSparseAG3DConv(k, X, Y, Z, fastconv, n, m, l, apply=None, _samples=None):
    """SparsifyingMatrix3D.jl:1659-1917: As*G truncated to the stencils (rows Values[c] * Entries[c])."""
    samples = _samples if _samples is not None else _sample_classes_3d(fastconv, n, m, l, apply)
    Indices, Values = entriesSparseA3D(k, X, Y, Z, fastconv, n, m, l, apply, samples)
    Entries = entriesSparseG3D(k, X, Y, Z, fastconv, n, m, l, apply, samples)
    ValuesAG = [np.asarray(v).reshape(1, -1) @ e for v, e in zip(Values, Entries)]
    return _assemble(n * m * l, _rows3(n, m, l), Indices, ValuesAG)


def sparsifying_matrices_3d(k, X, Y, Z, fastconv, n, m, l, nu, apply=None):
    """examples/example3D.jl:56-61 in one sampling pass: (As, Mapproxsp = As + k^2 AG diag(nu))."""
    samples = _sample_classes_3d(fastconv, n, m, l, apply)
    As = buildSparseA3DConv(k, X, Y, Z, fastconv, n, m, l, apply, samples)
    AG = buildSparseAG3DConv(k, X, Y, Z, fastconv, n, m, l, apply, samples)
    return As, _system_matrix(As, AG, k, nu)


# ------------------------------------------------------------------------------------------- 2-D
def _ind_relative_2d(n):
    # IndRelative of SparsifyingMatrix2D.jl:110-113 (rows: y offset -1,0,+1 ... as written upstream)
    return np.array([[-n - 1, -n, -n + 1], [-1, 0, 1], [n - 1, n, n + 1]], dtype=np.int64)


def _jl(a):
    """Julia's a[:] (column-major flattening)."""
    return np.asarray(a).reshape(-1, order="F")


def _centres_2d(n, m, strict):
    """Representative points of the interior and of the four edges (SparsifyingMatrix2D.jl:119,131,140,149,158).
    Upstream asserts an odd number of points (:106); strict=False takes the nearest interior point on even grids
    (any interior point gives the same stencil up to the phase ambiguity)."""
    N = n * m
    if n % 2 == 1 and m % 2 == 1:
        return [int(np.rint(v)) for v in (n * (m - 1) / 2 + (n + 1) / 2, n * (m - 1) / 2 + 1, n * (m - 1) / 2,
                                          (n + 1) / 2, N - (n + 1) / 2)]
    if strict:
        raise AssertionError("mod(length(X),2) == 1  (SparsifyingMatrix2D.jl:106)")
    jm, im = m // 2, n // 2
    return [im + 1 + n * jm, 1 + n * jm, n * jm, im + 1, N - n + im]


def sampleGConv(k, X, Y, indS, fastconv, apply=None):
    """FastConvolution.jl:278-306."""
    return _sample_rows(fastconv, len(X), indS, apply)


def _classes_2d(n, m, strict):
    """(stencil for As, stencil order used for G) per class, in upstream's order; the edge orderings of
    entriesSparseGConv (:293-304) differ from entriesSparseAConv's - kept."""
    IR = _ind_relative_2d(n)
    vol, fz1, fz2, fx1, fx2 = _centres_2d(n, m, strict)
    N = n * m
    c3, c4 = np.array([0, 1, -n, -n + 1]), np.array([0, -1, -n, -n - 1])
    return [
        (vol, _jl(IR), _jl(IR)),
        (fz1, _jl(IR[:, 1:3]), np.array([0, 1, n, n + 1, -n, -n + 1])),
        (fz2, _jl(IR[:, 0:2]), np.array([-1, 0, n, n - 1, -n, -n - 1])),
        (fx1, _jl(IR[1:3, :]), np.array([-1, 0, 1, n, n + 1, n - 1])),
        (fx2, _jl(IR[0:2, :]), np.array([-1, 0, 1, -n, -n + 1, -n - 1])),
        (1, _jl(IR[1:3, 1:3]), np.array([0, 1, n, n + 1])),
        (n, _jl(IR[1:3, 0:2]), np.array([0, -1, n, n - 1])),
        (N - n + 1, c3, c3),
        (N, c4, c4),
    ]


def _rows2(n, m):
    Ind = np.arange(1, n * m + 1, dtype=np.int64).reshape((n, m), order="F")
    return [_jl(Ind[1:-1, 1:-1]), _jl(Ind[0, 1:-1]), _jl(Ind[-1, 1:-1]), _jl(Ind[1:-1, 0]), _jl(Ind[1:-1, -1]),
            Ind[0, 0], Ind[-1, 0], Ind[0, -1], Ind[-1, -1]]


def _sample_classes_2d(fastconv, n, m, apply, strict):
    """One sampling pass over the 9 boundary classes: (relA, null vector, G block in entriesSparseGConv's stencil order).
    The edge orderings of entriesSparseGConv (:293-304) are permutations of entriesSparseAConv's stencils, so the block
    G[indG, indG] is the sampled block G[indA, indA] re-ordered - no second round of applies."""
    out = []
    sample = _sampler(fastconv, n * m, apply)
    for centre, relA, relG in _classes_2d(n, m, strict):
        relA = np.asarray(relA, dtype=np.int64)
        relG = np.asarray(relG, dtype=np.int64)
        pos = {int(r): i for i, r in enumerate(relA)}
        perm = np.array([pos[int(r)] for r in relG], dtype=np.int64)
        rows = sample(centre + relA)
        out.append((relA, rows.null_vector(), rows.block(perm)))
        rows.free()
    return out


def entriesSparseAConv(k, X, Y, fastconv, n, m, apply=None, strict=True, _samples=None):
    """SparsifyingMatrix2D.jl:104-201 -> (Indices, Entries)."""
    samples = _samples if _samples is not None else _sample_classes_2d(fastconv, n, m, apply, strict)
    return [rel for rel, v, blk in samples], [v for rel, v, blk in samples]


def entriesSparseGConv(k, X, Y, fastconv, n, m, apply=None, strict=True, _samples=None):
    """SparsifyingMatrix2D.jl:278-350."""
    samples = _samples if _samples is not None else _sample_classes_2d(fastconv, n, m, apply, strict)
    return [blk for rel, v, blk in samples]


def buildSparseAConv(k, X, Y, fastconv, n, m, apply=None, strict=True, _cache=None):
    """SparsifyingMatrix2D.jl:888-966."""
    Indices, Values = _cache if _cache is not None else entriesSparseAConv(k, X, Y, fastconv, n, m, apply, strict)
    return _assemble(n * m, _rows2(n, m), Indices, Values)


def buildSparseAGConv(k, X, Y, fastconv, n, m, apply=None, strict=True, _cache=None, _samples=None):
    """SparsifyingMatrix2D.jl:441-532: rows Values[c] * Entries[c]."""
    if _samples is not None:
        Indices, Values = entriesSparseAConv(k, X, Y, fastconv, n, m, apply, strict, _samples)
    else:
        Indices, Values = _cache if _cache is not None else entriesSparseAConv(k, X, Y, fastconv, n, m, apply, strict)
    Entries = entriesSparseGConv(k, X, Y, fastconv, n, m, apply, strict, _samples)
    # literal upstream: ValuesAG = Values[c] * Entries[c] (:456-532) - on the four edges Values follows entriesSparseAConv's
    # stencil order and Entries entriesSparseGConv's (:293-304); the mismatch is the reference's and is kept
    ValuesAG = [np.asarray(v).reshape(1, -1) @ e for v, e in zip(Values, Entries)]
    return _assemble(n * m, _rows2(n, m), Indices, ValuesAG)


def sparsifying_matrices_2d(k, X, Y, fastconv, n, m, nu, apply=None, strict=True):
    """examples/example.jl:64-67 with the *Conv builders, in one sampling pass: (As, Mapproxsp = As + k^2 AG diag(nu))."""
    samples = _sample_classes_2d(fastconv, n, m, apply, strict)
    As = buildSparseAConv(k, X, Y, fastconv, n, m, apply, strict, _cache=entriesSparseAConv(k, X, Y, fastconv, n, m, apply, strict, samples))
    AG = buildSparseAGConv(k, X, Y, fastconv, n, m, apply, strict, _samples=samples)
    return As, _system_matrix(As, AG, k, nu)
