"""Synthetic problem inputs for the bench and examples (host side, product code).

These build the *inputs* of the operators the way the reference's setup code does
(Green's spectrum of the Greengard-Vico truncated kernel, contrast nu); they are not part of
the GPU hot path.  The spectrum is even in every wave-number component, so only one
quadrant (octant) is evaluated and mirrored - bit-identical to evaluating everything,
because the squares kx^2 are bitwise equal for +-kx.
"""
from __future__ import annotations

import numpy as np
from scipy.special import hankel1, jv


def nu_gaussian_2d(X, Y):
    """examples/example.jl:48."""
    return 0.3 * np.exp(-40 * (X ** 2 + Y ** 2)) * (np.abs(X) < 0.48) * (np.abs(Y) < 0.48)


def gv_spectrum_2d(n, m, h, k):
    """GFFT of buildFastConvolution's Greengard_Vico branch (FastConvolution.jl:185-231;
    Gtruncated2D, Functions.jl:40-42), shape (4n, 4m), centred ordering."""
    Lp = 4.0 * (n * h)
    L = 1.5 * (n * h)
    kx = (2 * np.pi / Lp) * np.arange(-2 * n, 1, dtype=np.float64)     # -2n .. 0
    ky = (2 * np.pi / Lp) * np.arange(-2 * m, 1, dtype=np.float64)
    KX = np.repeat(kx[:, None], ky.size, axis=1)
    KY = np.repeat(ky[None, :], kx.size, axis=0)
    s = np.sqrt(KX ** 2 + KY ** 2)
    q = (1.0 + (1j * np.pi / 2 * L * hankel1(0, L * k)) * (s * jv(1, L * s))
         - (1j * np.pi / 2 * L * k * hankel1(1, L * k)) * jv(0, L * s)) / (s ** 2 - k ** 2)
    G = np.empty((4 * n, 4 * m), dtype=np.complex128)
    G[:2 * n + 1, :2 * m + 1] = q
    G[2 * n + 1:, :2 * m + 1] = q[2 * n - 1:0:-1, :]                  # kx = 1..2n-1  <- |kx|
    G[:, 2 * m + 1:] = G[:, 2 * m - 1:0:-1]
    return G


def gv_problem_2d(n, m=None, ppw=10.0, a=1.0, nu=nu_gaussian_2d):
    """Config C2: n (x m) grid, h = a/n, x = -a/2:h:a/2-h, k = 2 pi/(ppw h)."""
    m = n if m is None else m
    h = a / n
    x = -a / 2 + h * np.arange(n)
    y = -a / 2 * m / n + h * np.arange(m)
    k = 2 * np.pi / (ppw * h)
    X = np.repeat(x[:, None], m, axis=1).reshape(-1, order="F")
    Y = np.repeat(y[None, :], n, axis=0).reshape(-1, order="F")
    return np.asarray(nu(X, Y), dtype=np.float64), gv_spectrum_2d(n, m, h, k), k, h


def nu_gaussian_3d_grid(n, a=1.0):
    """examples/example3D.jl:43 on the n^3 grid x = -a/2:h:a/2-h (flattened, x fastest)."""
    h = a / n
    x = -a / 2 + h * np.arange(n)
    g = np.exp(-40 * x ** 2) * (np.abs(x) < 0.48)
    nu = 0.3 * g[:, None, None] * g[None, :, None] * g[None, None, :]
    return np.ascontiguousarray(nu.reshape(-1, order="F"))
