"""CPU check of the index logic of the Msp solve's batched matrix-vector kernel (csrc/msp_gemv.cuh).

The kernel body is written as __host__ __device__ phase functions; tests/msp_gemv_emu.cu runs exactly those functions CTA
by CTA on the CPU (nvcc builds it as a host library - no GPU involved).  Ragged batches in every addressing mode of the
solve (plain / index-gathered / assembled right-hand sides, scattered outputs) are compared with numpy for every
(LANES, UNR, XS) the tuner may pick.  The GPU parity of the whole solve is tests/test_gpu_msp.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "msp_gemv_emu.cu")
HDR = os.path.join(os.path.dirname(HERE), "fast_solver_lippmann_schwinger_b200", "csrc", "msp_gemv.cuh")
OUT = os.path.join(HERE, "_build", "libmsp_gemv_emu.so")


class Gemv2(C.Structure):
    _fields_ = [("M", C.c_void_p), ("moff", C.c_void_p), ("nrows", C.c_void_p), ("ncols", C.c_void_p),
                ("rows_p", C.c_int), ("cols_p", C.c_int), ("nodes", C.c_int),
                ("xmode", C.c_int), ("x", C.c_void_p), ("xstride", C.c_long), ("xidx", C.c_void_p), ("xg", C.c_void_p),
                ("f", C.c_void_p), ("sidx", C.c_void_p), ("pmap", C.c_void_p), ("tchild", C.c_void_p),
                ("Fp", C.c_int), ("Bpc", C.c_int),
                ("ymode", C.c_int), ("y0", C.c_void_p), ("y0stride", C.c_long), ("sign", C.c_double),
                ("out", C.c_void_p), ("ostride", C.c_long), ("oidx", C.c_void_p), ("og", C.c_void_p)]


@pytest.fixture(scope="module")
def emu():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available: the emulation library cannot be built")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run([nvcc, "-O1", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                        "-o", OUT, SRC], check=True, capture_output=True)
    L = C.CDLL(OUT)
    L.emu_msp_gemv2.restype = C.c_int
    L.emu_msp_gemv2.argtypes = [C.POINTER(Gemv2), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint), C.POINTER(C.c_long)]
    assert L.emu_sizeof_gemv2() == C.sizeof(Gemv2)
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def crandn(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


class Problem:
    """A ragged batch with the addressing of one launch of the solve."""

    def __init__(self, rng, nodes, rows_p, cols_p, xmode, ymode, scatter, sign, ragged=True, has_children=True):
        self.nodes, self.rows_p, self.cols_p = nodes, rows_p, cols_p
        lo_r, lo_c = (max(rows_p - 2, 0), max(cols_p - 2, 0)) if ragged else (rows_p, cols_p)
        self.nr = rng.integers(lo_r, rows_p + 1, nodes).astype(np.int32)
        self.nc = rng.integers(lo_c, cols_p + 1, nodes).astype(np.int32)
        if ragged and nodes > 2:
            self.nr[1] = 0 if rows_p > 0 else 0              # an empty node in the middle of the batch
        sizes = self.nr.astype(np.int64) * self.nc
        self.moff = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        self.M = crandn(rng, int(sizes.sum()) + 1)
        self.a = Gemv2()
        a = self.a
        a.M, a.moff, a.nrows, a.ncols = _p(self.M), _p(self.moff), _p(self.nr), _p(self.nc)
        a.rows_p, a.cols_p, a.nodes = rows_p, cols_p, nodes
        a.xmode, a.ymode, a.sign = xmode, ymode, sign
        NG = 4 * nodes * max(rows_p, cols_p, 1) + 7
        self.X = np.zeros((nodes, max(cols_p, 1)), complex)     # the values X(t, c) the kernel must see
        self.keep = []
        # children's update vectors and the parent maps (xmode 2 / ymode 2)
        self.Fp = cols_p + rows_p if (xmode == 2 or ymode == 2) else 0
        self.Bpc = 5 + rows_p
        self.tchild = crandn(rng, 2 * nodes * self.Bpc + 1)
        self.pmap = None
        if (xmode == 2 or ymode == 2) and has_children:
            pm = rng.integers(-1, self.Bpc, (nodes, 2, self.Fp)).astype(np.int32)
            pm[rng.random(pm.shape) < 0.4] = -1
            self.pmap = pm
        a.pmap, a.tchild, a.Fp, a.Bpc = _p(self.pmap), _p(self.tchild), self.Fp, self.Bpc

        def children(t, p):
            v = 0.0
            if self.pmap is not None:
                for side in (0, 1):
                    k = self.pmap[t, side, p]
                    if k >= 0:
                        v = v + self.tchild[(2 * t + side) * self.Bpc + k]
            return v

        if xmode == 0:
            self.xstride = cols_p + 3
            self.x = crandn(rng, nodes * self.xstride + 1)
            a.x, a.xstride = _p(self.x), self.xstride
            for t in range(nodes):
                self.X[t, :cols_p] = self.x[t * self.xstride:t * self.xstride + cols_p]
        elif xmode == 1:
            self.xg = crandn(rng, NG)
            self.xidx = rng.integers(0, NG, (nodes, max(cols_p, 1))).astype(np.int32)
            self.xidx[rng.random(self.xidx.shape) < 0.1] = -1
            a.xidx, a.xg = _p(self.xidx), _p(self.xg)
            for t in range(nodes):
                for c in range(cols_p):
                    self.X[t, c] = self.xg[self.xidx[t, c]] if self.xidx[t, c] >= 0 else 0.0
        else:
            self.f = crandn(rng, NG)
            self.sidx = rng.integers(0, NG, (nodes, max(cols_p, 1))).astype(np.int32)
            self.sidx[rng.random(self.sidx.shape) < 0.1] = -1
            a.f, a.sidx = _p(self.f), _p(self.sidx)
            for t in range(nodes):
                for c in range(cols_p):
                    v = self.f[self.sidx[t, c]] if self.sidx[t, c] >= 0 else 0.0
                    self.X[t, c] = v + children(t, c)
        self.Y0 = np.zeros((nodes, max(rows_p, 1)), complex)
        if ymode == 1:
            self.y0stride = rows_p + 2
            self.y0 = crandn(rng, nodes * self.y0stride + 1)
            a.y0, a.y0stride = _p(self.y0), self.y0stride
            for t in range(nodes):
                self.Y0[t, :rows_p] = self.y0[t * self.y0stride:t * self.y0stride + rows_p]
        elif ymode == 2:
            for t in range(nodes):
                for r in range(rows_p):
                    self.Y0[t, r] = children(t, cols_p + r)
        self.scatter = scatter
        if scatter:
            self.og = np.full(NG, 7.0 + 0j)
            perm = rng.permutation(NG)[:nodes * max(rows_p, 1)].astype(np.int32).reshape(nodes, max(rows_p, 1))
            perm[rng.random(perm.shape) < 0.1] = -1
            self.oidx = perm
            a.oidx, a.og = _p(self.oidx), _p(self.og)
        else:
            self.ostride = rows_p + 1
            self.out = np.full(nodes * self.ostride + 1, 7.0 + 0j)
            a.out, a.ostride = _p(self.out), self.ostride

    def expected(self):
        if self.scatter:
            ref = np.full(self.og.shape, 7.0 + 0j)
        else:
            ref = np.full(self.out.shape, 7.0 + 0j)
        for t in range(self.nodes):
            nr, nc = int(self.nr[t]), int(self.nc[t])
            A = self.M[self.moff[t]:self.moff[t] + nr * nc].reshape(nr, nc)
            val = self.Y0[t, :nr] + self.a.sign * (A @ self.X[t, :nc])
            for r in range(nr):
                if self.scatter:
                    if self.oidx[t, r] >= 0:
                        ref[self.oidx[t, r]] = val[r]
                else:
                    ref[t * self.ostride + r] = val[r]
        return ref

    def result(self):
        return self.og if self.scatter else self.out

    def reset(self):
        self.result()[:] = 7.0 + 0j


SHAPES = [  # nodes, rows_p, cols_p
    (37, 16, 16),      # leaf-like: many nodes per CTA
    (5, 9, 20),
    (3, 70, 33),       # several CTAs per node for narrow groups, one for wide ones
    (2, 300, 130),     # cpn > 1 everywhere
    (1, 1, 1),
    (4, 6, 0),         # root of the downward sweep: no columns
    (130, 4, 24),
]
MODES = [  # xmode, ymode, scatter, sign
    (0, 0, False, 1.0),    # z = Sinv g_S, unfused
    (2, 0, False, 1.0),    # z = Sinv g_S, gather fused
    (0, 1, False, -1.0),   # t = g_B - F_BS z, unfused
    (0, 2, False, -1.0),   # fused
    (1, 1, True, -1.0),    # u_S = z - Y u_B
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", MODES)
def test_emulated_kernel_matches_numpy(emu, shape, mode):
    nodes, rows_p, cols_p = shape
    xmode, ymode, scatter, sign = mode
    rng = np.random.default_rng(1000 * nodes + 10 * rows_p + cols_p + xmode)
    P = Problem(rng, nodes, rows_p, cols_p, xmode, ymode, scatter, sign)
    ref = P.expected()
    scale = max(1.0, np.abs(ref).max())
    geo = (C.c_uint * 3)()
    smem = C.c_long()
    first = {}
    for ll in range(6):
        for ul in range(4):
            for xs in (0, 1):
                P.reset()
                grid = emu.emu_msp_gemv2(C.byref(P.a), ll, ul, xs, geo, C.byref(smem))
                assert grid > 0
                lanes, unr = 1 << ll, 1 << ul
                rblocks, npc, cpn = geo[0], geo[1], geo[2]
                assert rblocks == max(1, -(-rows_p // unr)) and (npc == 1 or cpn == 1)
                assert npc * rblocks <= 256 // lanes or cpn > 1
                assert smem.value == (npc * cols_p * 16 if xs else 0)
                got = P.result()
                assert np.abs(got - ref).max() <= 1e-12 * scale, (ll, ul, xs)
                # for a given LANES the bits do not depend on UNR / XS
                if ll in first:
                    assert np.array_equal(got, first[ll]), (ll, ul, xs)
                else:
                    first[ll] = got.copy()


def test_leaf_level_without_children(emu):
    """xmode 2 / ymode 2 with no parent maps (leaf depth): g_S = f[Sidx], g_B = 0."""
    rng = np.random.default_rng(5)
    for xmode, ymode in ((2, 0), (0, 2)):
        P = Problem(rng, 21, 12, 9, xmode, ymode, False, -1.0, has_children=False)
        ref = P.expected()
        for ll, ul, xs in ((0, 0, 0), (2, 2, 1), (3, 3, 1), (5, 1, 0)):
            P.reset()
            assert emu.emu_msp_gemv2(C.byref(P.a), ll, ul, xs, None, None) > 0
            assert np.abs(P.result() - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_default_choice_rule(emu):
    out = (C.c_int * 3)()
    for cols, xmode, want in ((16, 1, (2, 2, 1)), (33, 0, (3, 2, 0)), (64, 2, (3, 2, 1)), (65, 0, (4, 2, 0)), (2561, 1, (5, 2, 1)), (0, 1, (2, 2, 1))):
        emu.emu_default_choice(cols, xmode, out)
        assert tuple(out) == want, (cols, xmode, tuple(out))
