"""CPU oracle for the Lippmann-Schwinger hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement, line by line, of the reference's Julia code for
the path named in BASELINE.json (north_star).  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing in ``fast_solver_lippmann_schwinger_b200/`` (the
product) imports it; the product fails loudly when ``libls_cuda.so`` is missing.

PARITY UNPINNED: the reference ships no golden vectors, no assertions and cannot be run
here (Julia, FFTW, UMFPACK, MKL are absent).  What pins this
oracle instead (tests/test_oracle.py):
  * the dense Green matrix of ``buildConvMatrix`` (FastConvolution.jl:497-513) equals the
    trapezoidal FFT apply to rounding;
  * the Greengard-Vico apply reproduces the direct quadrature sum to quadrature accuracy;
  * structural invariants of ``As`` (nnz, rows per class, bandwidth), ``As*G`` far field
    suppression;
  * GMRES restatement against an independent dense least-squares Krylov solve.

Third-party arithmetic not present under /root/reference (named, unpinned - the reference
has no Project.toml/Manifest.toml): FFTW.jl (``fft``/``ifft``/``fftshift``/``ifftshift`` ->
scipy.fft pocketfft), SpecialFunctions.jl (``hankelh1``/``besselj`` -> scipy.special, AMOS in
both), SparseArrays (CSC SpMV restated below), SuiteSparse UMFPACK (``lu`` -> SuperLU
``splu``), IterativeSolvers.jl (``gmres!`` restated in oracle/gmres_is.py).

Conventions: complex128, Julia column-major with x fastest, i.e. a grid
function ``u`` of shape (n, m) is stored as ``u.reshape(-1, order="F")``.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import scipy.fft as sfft
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.special import hankel1, jv

WORKERS = int(os.environ.get("LS_ORACLE_WORKERS", os.cpu_count() or 1))


def _fft(a):
    return sfft.fftn(a, workers=WORKERS)


def _ifft(a):
    return sfft.ifftn(a, workers=WORKERS)


# --------------------------------------------------------------------------------------
# Functions.jl
# --------------------------------------------------------------------------------------
def Gtruncated2D(L, k, s):
    """Functions.jl:40-42.  Truncated-kernel spectrum in 2-D (Vico-Greengard-Ferrando)."""
    s = np.asarray(s, dtype=np.float64)
    return (1.0
            + (1j * np.pi / 2 * L * hankel1(0, L * k)) * (s * jv(1, L * s))
            - (1j * np.pi / 2 * L * k * hankel1(1, L * k)) * jv(0, L * s)) / (s ** 2 - k ** 2)


def Gtruncated3D(L, k, s):
    """Functions.jl:45-51.  Julia's sinc(x) = sin(pi x)/(pi x) == numpy.sinc."""
    s = np.asarray(s, dtype=np.float64)
    return (-1.0 + np.exp(1j * L * k) * (np.cos(L * s) - (1j * k * L * np.sinc(L * s / np.pi)))) / (k ** 2 - s ** 2)


def createIndices(row, col, val):
    """Functions.jl:7-29: Row = kron(row, ones(nn)); Col = kron(ones(mm), col) + Row."""
    row = np.atleast_1d(np.asarray(row, dtype=np.int64))
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val, dtype=np.complex128)
    assert col.size == val.size
    nn, mm = col.size, row.size
    Row = np.kron(row, np.ones(nn, dtype=np.int64))
    Col = np.kron(np.ones(mm, dtype=np.int64), col) + Row
    Val = np.kron(np.ones(mm, dtype=np.int64), val)
    return Row, Col, Val


def referenceValsTrapRule():
    """FastConvolution.jl:407-415 (Duan-Rokhlin corrected trapezoidal rule table)."""
    x = 2.0 ** (-np.arange(6.0))
    w = np.array([1 - 0.892j, 1 - 1.35j, 1 - 1.79j, 1 - 2.23j, 1 - 2.67j, 1 - 3.11j])
    return x, w


# --------------------------------------------------------------------------------------
# FastConvolution.jl : operator object and the apply
# --------------------------------------------------------------------------------------
@dataclass
class FastM:
    """FastConvolution.jl:11-27."""
    GFFT: np.ndarray          # (ne, me) complex128
    nu: np.ndarray            # (n*m,) float64
    ne: int
    me: int
    n: int
    m: int
    omega: float
    quadRule: str = "trapezoidal"   # ctor default, FastConvolution.jl:24

    def size(self, dim=None):
        """FastConvolution.jl:31-37 (Q1: size(M) is a tuple of tuples)."""
        if dim is not None:
            return self.nu.shape[0]
        return (self.nu.shape, self.nu.shape)

    def eltype(self):
        """FastConvolution.jl:39-41."""
        return self.GFFT.dtype

    def __mul__(self, b):
        """FastConvolution.jl:43-48."""
        return fastconvolution(self, b)

    def mul_(self, Y, b):
        """FastConvolution.jl:50-54  (mul!: Y[:] = M*b)."""
        Y[:] = fastconvolution(self, b)
        return Y


def fastconvolution(M: FastM, b):
    """FastConvolution.jl:58-107:  b + omega^2 * crop(ifft2(GFFT .* fft2(pad(nu .* b))))."""
    b = np.asarray(b, dtype=np.complex128)
    if M.quadRule == "trapezoidal":                                            # :64-83
        BExt = np.zeros((M.ne, M.me), dtype=np.complex128)
        BExt[:M.n, :M.m] = (M.nu * b).reshape((M.n, M.m), order="F")
        BFft = _fft(BExt)
        BFft = M.GFFT * BFft
        BExt = _ifft(BFft)
        B = M.omega ** 2 * BExt[M.n - 1:2 * M.n - 1, M.m - 1:2 * M.m - 1]      # :82 (1-based n:2n-1)
    elif M.quadRule == "Greengard_Vico":                                       # :84-103
        BExt = np.zeros((M.ne, M.me), dtype=np.complex128)
        BExt[:M.n, :M.m] = (M.nu * b).reshape((M.n, M.m), order="F")
        BFft = sfft.fftshift(_fft(BExt))
        BFft = M.GFFT * BFft
        BExt = _ifft(sfft.ifftshift(BFft))
        B = M.omega ** 2 * BExt[:M.n, :M.m]                                    # :101
    else:
        raise ValueError("unknown quadRule %r" % (M.quadRule,))
    return b + B.reshape(-1, order="F")                                        # :106


def FFTconvolution(M: FastM, b):
    """FastConvolution.jl:110-154: the bare convolution.

    Quirks kept: Q2 (trapezoidal multiplies by nu, :122; Greengard_Vico does not, :141) and
    Q3 (pads to (ne, ne) and crops n in both dimensions -> square grids only).
    """
    b = np.asarray(b, dtype=np.complex128)
    if M.quadRule == "trapezoidal":
        indMiddle = int(np.rint(M.n))                                          # :115
        BExt = np.zeros((M.ne, M.ne), dtype=np.complex128)                     # :120
        BExt[:M.n, :M.m] = (M.nu * b).reshape((M.n, M.m), order="F")           # :122
        BFft = _fft(BExt)
        BFft = M.GFFT * BFft
        BExt = _ifft(BFft)
        B = BExt[indMiddle - 1:indMiddle - 1 + M.n, indMiddle - 1:indMiddle - 1 + M.n]   # :132
    elif M.quadRule == "Greengard_Vico":
        BExt = np.zeros((M.ne, M.ne), dtype=np.complex128)                     # :139
        BExt[:M.n, :M.n] = b.reshape((M.n, M.n), order="F")                    # :141
        BFft = sfft.fftshift(_fft(BExt))
        BFft = M.GFFT * BFft
        BExt = _ifft(sfft.ifftshift(BFft))
        B = BExt[:M.n, :M.n]                                                   # :151
    else:
        raise ValueError("unknown quadRule %r" % (M.quadRule,))
    return B.reshape(-1, order="F").copy()


def grid2d(x, y):
    """X = repeat(x,1,m)[:], Y = repeat(y',n,1)[:]   (example.jl:39-40)."""
    n, m = len(x), len(y)
    X = np.repeat(np.asarray(x)[:, None], m, axis=1).reshape(-1, order="F")
    Y = np.repeat(np.asarray(y)[None, :], n, axis=0).reshape(-1, order="F")
    return X, Y


def sampleGkernelpar(k, R, h):
    """FastConvolution.jl:348-401:  (i/4) h^2 H0^(1)(k R)."""
    return (1j / 4 * h ** 2) * hankel1(0, k * np.asarray(R, dtype=np.float64))


def buildGConv(x, y, h, n, m, D0, k):
    """FastConvolution.jl:425-469 (odd n only; the grid must be centred on the origin, Q6)."""
    if n % 2 != 1:
        raise ValueError("so far only works for n odd (FastConvolution.jl:441)")
    xe = x[0] - (n - 1) / 2 * h + h * np.arange(2 * n - 1)
    ye = y[0] - (m - 1) / 2 * h + h * np.arange(2 * m - 1)
    Xe = np.repeat(xe[:, None], 2 * m - 1, axis=1)
    Ye = np.repeat(ye[None, :], 2 * n - 1, axis=0)
    R = np.sqrt(Xe ** 2 + Ye ** 2)
    # findall(R.==0)[1] (:456) needs an exact zero; a centred grid built by range
    # arithmetic gives one in Julia - we take the minimum to be robust to 1-ulp noise.
    idx = np.unravel_index(np.argmin(R), R.shape)
    R[idx] = 1.0
    Ge = sampleGkernelpar(k, R, h)
    Ge[idx] = 1j / 4 * D0 * h ** 2
    return Ge


def gv_spectrum_2d(n, m, h, k):
    """The Greengard_Vico branch of buildFastConvolution, FastConvolution.jl:185-231.

    Lp = 4 n h, L = 1.5 n h (:187-188), centred wave numbers kx = -2n:2n-1 (:194-195).
    Returns GFFT of shape (4n, 4m) in the reference's (centred) ordering.
    """
    Lp = 4.0 * (n * h)
    L = 1.5 * (n * h)
    kx = np.arange(-2 * n, 2 * n, dtype=np.float64)
    ky = np.arange(-2 * m, 2 * m, dtype=np.float64)
    KX = (2 * np.pi / Lp) * np.repeat(kx[:, None], 4 * m, axis=1)
    KY = (2 * np.pi / Lp) * np.repeat(ky[None, :], 4 * n, axis=0)
    S = np.sqrt(KX ** 2 + KY ** 2)
    return Gtruncated2D(L, k, S)


def buildFastConvolution(x, y, h, k, nu, quadRule="trapezoidal"):
    """FastConvolution.jl:170-236.  ``nu`` is a callable nu(X, Y) on flattened grids."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n, m = len(x), len(y)
    X, Y = grid2d(x, y)
    if quadRule == "trapezoidal":
        _, D = referenceValsTrapRule()
        D0 = D[int(np.rint(k * h)) - 1]                                         # :176 (1-based)
        Ge = buildGConv(x, y, h, n, m, D0, k)
        GFFT = _fft(Ge)
        return FastM(GFFT, np.asarray(nu(X, Y), dtype=np.float64), 2 * n - 1, 2 * m - 1, n, m, k)
    if quadRule == "Greengard_Vico":
        # (abs(x[end]-x[1]) + h) == n*h for a uniform grid
        GFFT = gv_spectrum_2d(n, m, h, k)
        return FastM(GFFT, np.asarray(nu(X, Y), dtype=np.float64), 4 * n, 4 * m, n, m, k,
                     quadRule="Greengard_Vico")
    raise ValueError(quadRule)


def buildConvMatrix(k, X, Y, D0, h):
    """FastConvolution.jl:497-513: dense N x N Green matrix (small-n cross-check only)."""
    N = len(X)
    G = np.zeros((N, N), dtype=np.complex128)
    for ii in range(N):
        r = np.sqrt((X - X[ii]) ** 2 + (Y - Y[ii]) ** 2)
        r[ii] = 1.0
        G[ii, :] = 1j / 4 * hankel1(0, k * r) * h ** 2
        G[ii, ii] = 1j / 4 * D0 * h ** 2
    return G


# --------------------------------------------------------------------------------------
# FastConvolution3D.jl
# --------------------------------------------------------------------------------------
@dataclass
class FastM3D:
    """FastConvolution3D.jl:7-26."""
    GFFT: np.ndarray          # (ne, me, le)
    nu: np.ndarray            # (n*m*l,)
    ne: int
    me: int
    le: int
    n: int
    m: int
    l: int
    omega: float
    quadRule: str = "Greengard_Vico"

    def __mul__(self, b):
        """FastConvolution3D.jl:31-37."""
        b = np.asarray(b, dtype=np.complex128)
        B = self.omega ** 2 * FFTconvolution3D(self, self.nu * b)
        return b + B


def FFTconvolution3D(M: FastM3D, b):
    """FastConvolution3D.jl:39-63 (Q3: pads (ne, ne, le))."""
    b = np.asarray(b, dtype=np.complex128)
    BExt = np.zeros((M.ne, M.ne, M.le), dtype=np.complex128)                   # :48
    BExt[:M.n, :M.m, :M.l] = b.reshape((M.n, M.m, M.l), order="F")             # :50
    BFft = sfft.fftshift(_fft(BExt))
    BFft *= M.GFFT
    BExt = _ifft(sfft.ifftshift(BFft))
    return BExt[:M.n, :M.m, :M.l].reshape(-1, order="F").copy()


def gv_spectrum_3d(n, m, l, h, k):
    """Even-n branch of buildFastConvolution3D, FastConvolution3D.jl:72-101."""
    Lp = 4.0 * (n * h)
    L = 1.8 * (n * h)
    kx = (2 * np.pi / Lp) * np.arange(-2 * n, 2 * n, dtype=np.float64)
    ky = (2 * np.pi / Lp) * np.arange(-2 * m, 2 * m, dtype=np.float64)
    kz = (2 * np.pi / Lp) * np.arange(-2 * l, 2 * l, dtype=np.float64)
    S = np.sqrt(kx[:, None, None] ** 2 + ky[None, :, None] ** 2 + kz[None, None, :] ** 2)
    return Gtruncated3D(L, k, S)


def grid3d(x, y, z):
    """example3D.jl:33-39."""
    n, m, l = len(x), len(y), len(z)
    X = np.broadcast_to(np.asarray(x)[:, None, None], (n, m, l)).reshape(-1, order="F")
    Y = np.broadcast_to(np.asarray(y)[None, :, None], (n, m, l)).reshape(-1, order="F")
    Z = np.broadcast_to(np.asarray(z)[None, None, :], (n, m, l)).reshape(-1, order="F")
    return X, Y, Z


def buildFastConvolution3D(x, y, z, h, k, nu):
    """FastConvolution3D.jl:68-132, even n only (the odd branch is unfinished upstream)."""
    n, m, l = len(x), len(y), len(z)
    if n % 2 != 0:
        raise ValueError("odd-n 3-D branch (FastConvolution3D.jl:102-128) is not targeted")
    X, Y, Z = grid3d(x, y, z)
    GFFT = gv_spectrum_3d(n, m, l, h, k)
    return FastM3D(GFFT, np.asarray(nu(X, Y, Z), dtype=np.float64), 4 * n, 4 * m, 4 * l, n, m, l, k)


# --------------------------------------------------------------------------------------
# Sparsifying matrices (SparsifyingMatrix2D.jl) - setup code, restated only to obtain
# realistic As / Msp inputs for the SpMV and the preconditioned GMRES.
# --------------------------------------------------------------------------------------
def sampleG(k, X, Y, indS, D0):
    """FastConvolution.jl:239-275.  ``indS`` is 1-based (Julia)."""
    h = abs(X[1] - X[0])
    indS = np.asarray(indS, dtype=np.int64)
    R = np.empty((len(indS), len(X)))
    for i, ii in enumerate(indS):
        R[i, :] = np.sqrt((X - X[ii - 1]) ** 2 + (Y - Y[ii - 1]) ** 2)
        R[i, ii - 1] = 1.0
    Gc = sampleGkernelpar(k, R, h)
    for i, ii in enumerate(indS):
        Gc[i, ii - 1] = 1j / 4 * D0 * h ** 2
    return Gc


def _ind_relative(n):
    """IndRelative (SparsifyingMatrix2D.jl:11-14) as a 3x3 array (row-major literal)."""
    return np.array([[-n - 1, -n, -n + 1],
                     [-1, 0, 1],
                     [n - 1, n, n + 1]], dtype=np.int64)


def _jl(a):
    """Julia's A[:] - column-major flatten."""
    return np.asarray(a).reshape(-1, order="F")


def _centres(n, m, strict=True):
    """1-based stencil-centre indices used by entriesSparseA / entriesSparseG.

    With odd n*m these are the reference's expressions (SparsifyingMatrix2D.jl:20,32,40,48,
    56,64-67).  The reference asserts odd N (:7); for even sizes (GPU power-of-two configs)
    ``strict=False`` relaxes the centre to a valid interior / edge-midpoint index - any such
    point yields the same stencil up to the SVD phase.
    """
    N = n * m
    if N % 2 == 1:
        vol = n * (m - 1) / 2 + (n + 1) / 2
        fz1 = n * (m - 1) / 2 + 1
        fz2 = n * (m - 1) / 2
        fx1 = (n + 1) / 2
        fx2 = N - (n + 1) / 2
        c = [int(np.rint(v)) for v in (vol, fz1, fz2, fx1, fx2)]
    else:
        if strict:
            raise AssertionError("mod(length(X),2) == 1  (SparsifyingMatrix2D.jl:7)")
        jm, im = m // 2, n // 2          # 0-based middle row / column
        vol = im + 1 + n * jm
        fz1 = 1 + n * jm                  # x = xmin edge, middle in y
        fz2 = n + n * jm - n              # x = xmax edge: index of (n, jm)  (= n*jm in 1-based)
        fx1 = im + 1                      # y = ymin edge
        fx2 = N - n + im + 1 - 1          # y = ymax edge, mirrors N-(n+1)/2
        c = [vol, fz1, fz2, fx1, fx2]
    return c


def sampleGConv(k, X, Y, indS, fastconv: FastM):
    """FastConvolution.jl:278-306: rows of the discrete operator by FFTconvolution of unit vectors (1-based indS)."""
    indS = np.asarray(indS, dtype=np.int64)
    G = np.zeros((len(indS), len(X)), dtype=np.complex128)
    for i, s0 in enumerate(indS):
        e = np.zeros(len(X), dtype=np.complex128)
        e[s0 - 1] = 1.0
        G[i, :] = FFTconvolution(fastconv, e)
    return G


def entriesSparseA(k, X, Y, D0, n, m, strict=True, _sampler=None):
    """SparsifyingMatrix2D.jl:5-102: one stencil (row vector) per boundary class.
    (_sampler replaces sampleG: entriesSparseAConv, :104-201, is the same code with sampleGConv.)"""
    sampleG = _sampler if _sampler is not None else globals()["sampleG"]
    IR = _ind_relative(n)
    N = n * m
    vol, fz1, fz2, fx1, fx2 = _centres(n, m, strict)
    allidx = np.arange(1, N + 1)
    Entries, Indices = [], []

    def one_class(ind, rel):
        indC = np.setdiff1d(allidx, ind)
        GS = sampleG(k, X, Y, ind, D0)[:, indC - 1]
        U, s, Vh = np.linalg.svd(GS, full_matrices=False)
        Entries.append(np.conj(U[:, -1]))          # U[:,end]'  (adjoint -> conjugated row)
        Indices.append(np.asarray(rel, dtype=np.int64))

    one_class(vol + _jl(IR), _jl(IR))                                   # interior  :20-28
    one_class(fz1 + _jl(IR[:, 1:3]), _jl(IR[:, 1:3]))                   # x = xmin  :32-37
    one_class(fz2 + _jl(IR[:, 0:2]), _jl(IR[:, 0:2]))                   # x = xmax  :40-45
    one_class(fx1 + _jl(IR[1:3, :]), _jl(IR[1:3, :]))                   # y = ymin  :48-53
    one_class(fx2 + _jl(IR[0:2, :]), _jl(IR[0:2, :]))                   # y = ymax  :56-61
    one_class(1 + _jl(IR[1:3, 1:3]), _jl(IR[1:3, 1:3]))                 # corners   :64-78
    one_class(n + _jl(IR[1:3, 0:2]), _jl(IR[1:3, 0:2]))                 #           :81-86
    one_class(n * m - n + 1 + np.array([0, 1, -n, -n + 1]), [0, 1, -n, -n + 1])   # :89-94
    one_class(n * m + np.array([0, -1, -n, -n - 1]), [0, -1, -n, -n - 1])         # :97-102
    return Indices, Entries


def entriesSparseG(k, X, Y, D0, n, m, strict=True, _sampler=None):
    """SparsifyingMatrix2D.jl:205-275: G restricted to each class's stencil.

    The edge orderings here differ from entriesSparseA's (a reference quirk, kept).
    (_sampler: entriesSparseGConv, :278-350, is the same code with sampleGConv.)
    """
    sampleG = _sampler if _sampler is not None else globals()["sampleG"]
    IR = _ind_relative(n)
    N = n * m
    vol, fz1, fz2, fx1, fx2 = _centres(n, m, strict)
    out = []

    def one(ind):
        ind = np.asarray(ind, dtype=np.int64)
        out.append(sampleG(k, X, Y, ind, D0)[:, ind - 1])

    one(vol + _jl(IR))
    one(fz1 + np.array([0, 1, n, n + 1, -n, -n + 1]))
    one(fz2 + np.array([-1, 0, n, n - 1, -n, -n - 1]))
    one(fx1 + np.array([-1, 0, 1, n, n + 1, n - 1]))
    one(fx2 + np.array([-1, 0, 1, -n, -n + 1, -n - 1]))
    one(1 + np.array([0, 1, n, n + 1]))
    one(n + np.array([0, -1, n, n - 1]))
    one(N - n + 1 + np.array([0, 1, -n, -n + 1]))
    one(N + np.array([0, -1, -n, -n - 1]))
    return out


def _class_rows(n, m):
    """Row index sets in the order buildSparseA uses them (SparsifyingMatrix2D.jl:820-875)."""
    Ind = np.arange(1, n * m + 1, dtype=np.int64).reshape((n, m), order="F")
    return [_jl(Ind[1:-1, 1:-1]), _jl(Ind[0, 1:-1]), _jl(Ind[-1, 1:-1]), _jl(Ind[1:-1, 0]),
            _jl(Ind[1:-1, -1]), Ind[0, 0], Ind[-1, 0], Ind[0, -1], Ind[-1, -1]]


def _assemble(n, m, Indices, Values):
    rows, cols, vals = [], [], []
    for rset, ind, val in zip(_class_rows(n, m), Indices, Values):
        R, C, V = createIndices(rset, ind, np.asarray(val).reshape(-1))
        rows.append(R)
        cols.append(C)
        vals.append(V)
    R = np.concatenate(rows) - 1
    C = np.concatenate(cols) - 1
    V = np.concatenate(vals)
    A = sp.coo_matrix((V, (R, C)), shape=(n * m, n * m)).tocsc()    # sparse(rowA,colA,valA)
    A.sum_duplicates()
    A.sort_indices()
    return A


def buildSparseA(k, X, Y, D0, n, m, strict=True, _cache=None):
    """SparsifyingMatrix2D.jl:806-884.  Returns scipy CSC (same storage as SparseMatrixCSC)."""
    Indices, Values = _cache if _cache is not None else entriesSparseA(k, X, Y, D0, n, m, strict)
    return _assemble(n, m, Indices, Values)


def buildSparseAG(k, X, Y, D0, n, m, strict=True, _cache=None):
    """SparsifyingMatrix2D.jl:351-438:  rows Values[c] * Entries[c]."""
    Indices, Values = _cache if _cache is not None else entriesSparseA(k, X, Y, D0, n, m, strict)
    Entries = entriesSparseG(k, X, Y, D0, n, m, strict)
    ValuesAG = [np.asarray(v).reshape(1, -1) @ e for v, e in zip(Values, Entries)]
    return _assemble(n, m, Indices, ValuesAG)


# --------------------------------------------------------------------------------------
# 3-D sparsifying matrices (SparsifyingMatrix3D.jl) - setup code of examples/example3D.jl:56-61,
# restated to obtain a realistic 27-point As / Msp for the 3-D SpMV and preconditioned GMRES.
# --------------------------------------------------------------------------------------
# Boundary classes in the order entriesSparseA3D pushes them (SparsifyingMatrix3D.jl:1136-1408) and
# buildSparseA3DConv consumes them (:1410-1653): interior, 6 faces, 12 edges ("vertices" upstream), 8 corners.
# Per dimension: "lo" = first grid plane (stencil offsets 0,+1), "mid" = interior (-1,0,+1), "hi" = last (-1,0).
_CLASSES_3D = [
    ("mid", "mid", "mid"),
    ("lo", "mid", "mid"), ("hi", "mid", "mid"), ("mid", "lo", "mid"), ("mid", "hi", "mid"),
    ("mid", "mid", "lo"), ("mid", "mid", "hi"),
    ("lo", "lo", "mid"), ("hi", "lo", "mid"), ("lo", "hi", "mid"), ("hi", "hi", "mid"),
    ("lo", "mid", "lo"), ("hi", "mid", "lo"), ("lo", "mid", "hi"), ("hi", "mid", "hi"),
    ("mid", "lo", "lo"), ("mid", "hi", "lo"), ("mid", "lo", "hi"), ("mid", "hi", "hi"),
    ("lo", "lo", "lo"), ("hi", "lo", "lo"), ("lo", "hi", "lo"), ("hi", "hi", "lo"),
    ("lo", "lo", "hi"), ("hi", "lo", "hi"), ("lo", "hi", "hi"), ("hi", "hi", "hi"),
]
_OFFS_3D = {"lo": (0, 1), "mid": (-1, 0, 1), "hi": (-1, 0)}


def _stencil_offsets_3d(cls, n, m):
    """Ind_relative[...][:] of SparsifyingMatrix3D.jl:1147-1158 restricted to the class (x fastest)."""
    cx, cy, cz = cls
    return np.array([dx + n * dy + n * m * dz for dz in _OFFS_3D[cz] for dy in _OFFS_3D[cy] for dx in _OFFS_3D[cx]],
                    dtype=np.int64)


def _class_centre_3d(cls, n, m, l):
    """1-based index of the representative point: plane 1 / round(n/2) / n per dimension (changeInd3D, :7-11)."""
    def coord(c, nn):
        return 1 if c == "lo" else (nn if c == "hi" else int(round(nn / 2)))     # Julia round: half to even, as Python
    i, j, p = coord(cls[0], n), coord(cls[1], m), coord(cls[2], l)
    return (p - 1) * n * m + (j - 1) * n + i


def sampleG3D(k, X, Y, Z, indS, fastconv: FastM3D, toeplitz=True):
    """FastConvolution3D.jl:136-160: row i = FFTconvolution(fastconv, e_{indS[i]}) (1-based indS).

    toeplitz=False runs the applies literally.  toeplitz=True reads the same numbers from the spatial kernel
    g = ifftn(ifftshift(GFFT)) (the apply is a circular convolution with g on the padded grid, cropped), which is
    what makes sampling 343 rows affordable; tests/test_oracle.py checks both paths against each other."""
    indS = np.asarray(indS, dtype=np.int64)
    N = fastconv.n * fastconv.m * fastconv.l
    if not toeplitz:
        G = np.zeros((len(indS), N), dtype=np.complex128)
        for i, s0 in enumerate(indS):
            e = np.zeros(N, dtype=np.complex128)
            e[s0 - 1] = 1.0
            G[i, :] = FFTconvolution3D(fastconv, e)
        return G
    g = getattr(fastconv, "_g_spatial", None)
    if g is None:
        g = _ifft(sfft.ifftshift(fastconv.GFFT))
        object.__setattr__(fastconv, "_g_spatial", g)
    n, m, l = fastconv.n, fastconv.m, fastconv.l
    I = np.arange(n)[:, None, None]
    J = np.arange(m)[None, :, None]
    P = np.arange(l)[None, None, :]
    G = np.empty((len(indS), N), dtype=np.complex128)
    for i, s0 in enumerate(indS - 1):
        si, sj, sp_ = s0 % n, (s0 // n) % m, s0 // (n * m)
        G[i, :] = g[(I - si) % fastconv.ne, (J - sj) % fastconv.ne, (P - sp_) % fastconv.le].reshape(-1, order="F")
    return G


def entriesSparseA3D(k, X, Y, Z, fastconv: FastM3D, n, m, l, toeplitz=True):
    """SparsifyingMatrix3D.jl:1136-1408: per boundary class the last left singular vector of G sampled from the
    class's stencil points to every other grid point.  Returns (Indices, Entries) like upstream."""
    N = n * m * l
    allidx = np.arange(1, N + 1)
    Indices, Entries = [], []
    for cls in _CLASSES_3D:
        rel = _stencil_offsets_3d(cls, n, m)
        ind = _class_centre_3d(cls, n, m, l) + rel
        indC = np.setdiff1d(allidx, ind)
        GS = sampleG3D(k, X, Y, Z, ind, fastconv, toeplitz)[:, indC - 1]
        U, s, Vh = np.linalg.svd(GS, full_matrices=False)
        Entries.append(np.conj(U[:, -1]))            # U[:,end]'
        Indices.append(rel)
    return Indices, Entries


def entriesSparseG3D(k, X, Y, Z, fastconv: FastM3D, n, m, l, toeplitz=True):
    """SparsifyingMatrix3D.jl:963-1135: G restricted to each class's own stencil (same class order)."""
    out = []
    for cls in _CLASSES_3D:
        ind = _class_centre_3d(cls, n, m, l) + _stencil_offsets_3d(cls, n, m)
        out.append(sampleG3D(k, X, Y, Z, ind, fastconv, toeplitz)[:, ind - 1])
    return out


def _class_rows_3d(n, m, l):
    """Row sets Ind[...][:] in the order of buildSparseA3DConv (SparsifyingMatrix3D.jl:1427-1648)."""
    Ind = np.arange(1, n * m * l + 1, dtype=np.int64).reshape((n, m, l), order="F")
    sel = {"lo": slice(0, 1), "mid": slice(1, -1), "hi": slice(-1, None)}
    return [Ind[sel[cx], sel[cy], sel[cz]].reshape(-1, order="F") for cx, cy, cz in _CLASSES_3D]


def _assemble_3d(n, m, l, Indices, Values):
    rows, cols, vals = [], [], []
    for rset, ind, val in zip(_class_rows_3d(n, m, l), Indices, Values):
        R, C, V = createIndices(rset, ind, np.asarray(val).reshape(-1))
        rows.append(R)
        cols.append(C)
        vals.append(V)
    N = n * m * l
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows) - 1, np.concatenate(cols) - 1)), shape=(N, N)).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    return A


def buildSparseA3DConv(k, X, Y, Z, fastconv: FastM3D, n, m, l, _cache=None):
    """SparsifyingMatrix3D.jl:1410-1653 (method = "normal")."""
    Indices, Values = _cache if _cache is not None else entriesSparseA3D(k, X, Y, Z, fastconv, n, m, l)
    return _assemble_3d(n, m, l, Indices, Values)


def buildSparseAG3DConv(k, X, Y, Z, fastconv: FastM3D, n, m, l, _cache=None):
    """SparsifyingMatrix3D.jl:1659-1917: rows Values[c] * Entries[c] on the same stencils (As*G truncated)."""
    Indices, Values = _cache if _cache is not None else entriesSparseA3D(k, X, Y, Z, fastconv, n, m, l)
    Entries = entriesSparseG3D(k, X, Y, Z, fastconv, n, m, l)
    ValuesAG = [np.asarray(v).reshape(1, -1) @ e for v, e in zip(Values, Entries)]
    return _assemble_3d(n, m, l, Indices, ValuesAG)


def entriesSparseAConv(k, X, Y, fastconv: FastM, n, m, strict=True):
    """SparsifyingMatrix2D.jl:104-201."""
    return entriesSparseA(k, X, Y, None, n, m, strict, _sampler=lambda k_, X_, Y_, ind, D0: sampleGConv(k_, X_, Y_, ind, fastconv))


def buildSparseAConv(k, X, Y, fastconv: FastM, n, m, strict=True, _cache=None):
    """SparsifyingMatrix2D.jl:888-966."""
    Indices, Values = _cache if _cache is not None else entriesSparseAConv(k, X, Y, fastconv, n, m, strict)
    return _assemble(n, m, Indices, Values)


def buildSparseAGConv(k, X, Y, fastconv: FastM, n, m, strict=True, _cache=None):
    """SparsifyingMatrix2D.jl:441-532 with entriesSparseGConv (:278-350)."""
    Indices, Values = _cache if _cache is not None else entriesSparseAConv(k, X, Y, fastconv, n, m, strict)
    Entries = entriesSparseG(k, X, Y, None, n, m, strict, _sampler=lambda k_, X_, Y_, ind, D0: sampleGConv(k_, X_, Y_, ind, fastconv))
    ValuesAG = [np.asarray(v).reshape(1, -1) @ e for v, e in zip(Values, Entries)]
    return _assemble(n, m, Indices, ValuesAG)


def csc_matvec(A: sp.csc_matrix, x):
    """SparseArrays' ``A*x`` for CSC - the column-scatter loop (== sparseblas.jl:14-25 with
    alpha=1, beta=0).  scipy's csc_matvec runs the same loop in C."""
    return A @ np.asarray(x, dtype=np.complex128)


def csc_matvec_loops(colptr, rowval, nzval, x, nrows):
    """Pure-Python statement of the same loop on Julia's 1-based arrays (small cases)."""
    y = np.zeros(nrows, dtype=np.complex128)
    for col in range(len(colptr) - 1):
        for j in range(colptr[col], colptr[col + 1]):
            y[rowval[j - 1] - 1] += nzval[j - 1] * x[col]
    return y


def julia_csc_arrays(A: sp.csc_matrix):
    """(colptr, rowval, nzval) exactly as Julia's SparseMatrixCSC{ComplexF64,Int64} holds them."""
    A = A.tocsc()
    A.sort_indices()
    return (A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1,
            A.data.astype(np.complex128))


class SparsifyingPreconditioner:
    """preconditioner.jl:27-58 (UMFPACK branch; SuperLU stands in for UMFPACK)."""

    def __init__(self, Msp, As):
        self.Msp = Msp.tocsc()
        self.As = As.tocsc()
        self.MspInv = spla.splu(self.Msp)
        self.solverType = "UMFPACK"

    def solve(self, b):
        """``\\``  preconditioner.jl:132-145:  MspInv \\ (As*b)."""
        return self.MspInv.solve(csc_matvec(self.As, b))

    def ldiv_(self, b):
        """ldiv!  preconditioner.jl:147-166 (in place)."""
        b[:] = self.solve(b)
        return b


# --------------------------------------------------------------------------------------
# Synthetic problems used by tests and bench
# --------------------------------------------------------------------------------------
def nu_gaussian_2d(X, Y):
    """examples/example.jl:48."""
    return 0.3 * np.exp(-40 * (X ** 2 + Y ** 2)) * (np.abs(X) < 0.48) * (np.abs(Y) < 0.48)


def nu_gaussian_3d(X, Y, Z):
    """examples/example3D.jl:43."""
    return (0.3 * np.exp(-40 * (X ** 2 + Y ** 2 + Z ** 2)) * (np.abs(X) < 0.48)
            * (np.abs(Y) < 0.48) * (np.abs(Z) < 0.48))


def nu_plasma_2d(X, Y):
    """tests/plasma_example.jl:53-68 (discontinuous plasma profile)."""
    C = 0.4987

    def phi(x, y):
        return 1 - (x - 0.05 * (1 - x ** 2)) ** 2 - C * ((1 + 0.3 * x) ** 2) * y ** 2

    aa = [0.45, 0.196, 0.51, 0.195, 0.63]
    xI = [0.4, 0.54, -0.14, -0.5, 0.18]
    yI = [0, -0.28, 0.70, -0.01, 0.8]

    def g(x, y):
        return sum(a * np.exp(-((x - xi) ** 2 + (y - yi) ** 2) / 0.01) for a, xi, yi in zip(aa, xI, yI))

    def nu2(x, y):
        p = phi(x, y)
        return (p > 0.05) * (-1.5 * (p - 0.05) - g(x, y) * np.cos(0.9 * y))

    return -nu2(3 * X, 3 * Y)


def example_problem_2d(h=0.005, a=1.0, quadRule="Greengard_Vico", nu=nu_gaussian_2d, k=None):
    """examples/example.jl:30-54: x = -a/2:h:a/2 (odd n), k = 1/h."""
    npts = int(round(a / h)) + 1
    x = -a / 2 + h * np.arange(npts)
    k = 1.0 / h if k is None else k
    return x, buildFastConvolution(x, x, h, k, nu, quadRule=quadRule)


def pow2_problem_2d(n, ppw=10.0, nu=nu_gaussian_2d, a=1.0):
    """Config C2/C3: n = m power of two, h = a/n, x = -a/2:h:a/2-h, k = 2 pi/(ppw h)."""
    h = a / n
    x = -a / 2 + h * np.arange(n)
    k = 2 * np.pi / (ppw * h)
    return x, h, k, buildFastConvolution(x, x, h, k, nu, quadRule="Greengard_Vico")


def example_problem_3d(n, l=None, a=1.0, ppw=None, nu=nu_gaussian_3d):
    """examples/example3D.jl:18-61 on an n x n x l grid (the script ships h = 1/48, x = -a/2:h:a/2-h, i.e. n = l = 48,
    k = 1/h): operator, As, Msp = As + k^2 AG diag(nu) (:56-61) and the SparsifyingPreconditioner (SuperLU stands in
    for MKL PARDISO).  ppw overrides k = 1/h with k = 2 pi / (ppw h)."""
    l = n if l is None else l
    h = a / n
    x = -a / 2 + h * np.arange(n)
    z = -a / 2 * l / n + h * np.arange(l)
    k = (1.0 / h) if ppw is None else 2 * np.pi / (ppw * h)
    M = buildFastConvolution3D(x, x, z, h, k, nu)
    X, Y, Z = grid3d(x, x, z)
    cache = entriesSparseA3D(k, X, Y, Z, M, n, n, l)
    As = buildSparseA3DConv(k, X, Y, Z, M, n, n, l, _cache=cache)
    AG = buildSparseAG3DConv(k, X, Y, Z, M, n, n, l, _cache=cache)
    Msp = (As + k ** 2 * (AG @ sp.diags(M.nu))).tocsc()
    return (x, z), h, k, M, As, Msp, SparsifyingPreconditioner(Msp, As)


def pow2_problem_3d(n, ppw=10.0, nu=nu_gaussian_3d, a=1.0):
    """Config C4: n = m = l, h = a/n, k = 2 pi/(ppw h)."""
    h = a / n
    x = -a / 2 + h * np.arange(n)
    k = 2 * np.pi / (ppw * h)
    return x, h, k, buildFastConvolution3D(x, x, x, h, k, nu)
