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


def nu_plasma_2d(X, Y):
    """tests/plasma_example.jl:53-68: the discontinuous plasma profile nu(x, y) = -nu2(3x, 3y) (config 3)."""
    x, y = 3.0 * np.asarray(X, dtype=np.float64), 3.0 * np.asarray(Y, dtype=np.float64)
    phi = 1 - (x - 0.05 * (1 - x ** 2)) ** 2 - 0.4987 * ((1 + 0.3 * x) ** 2) * y ** 2
    bumps = ((0.45, 0.4, 0.0), (0.196, 0.54, -0.28), (0.51, -0.14, 0.70), (0.195, -0.5, -0.01), (0.63, 0.18, 0.8))
    g = sum(a * np.exp(-((x - xi) ** 2 + (y - yi) ** 2) / 0.01) for a, xi, yi in bumps)
    nu2 = (phi > 0.05) * (-1.5 * (phi - 0.05) - g * np.cos(0.9 * y))
    return -nu2


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


def nu_layered_3d_slab(n, p0, p1, a=1.0, values=(0.05, 0.10, 0.02, 0.08)):
    """Config 5's synthetic contrast ("layered, piecewise constant in z"; the reference's
    example3D_Polarized_traces.jl:45 itself ships the Gaussian bump): len(values) horizontal layers inside the box
    |x|, |y|, |z| < 0.48 a, zero outside, on the n^3 grid x = -a/2:h:a/2-h.  Returns the z planes [p0, p1) flattened
    (x fastest) - a rank's slab of the sharded operator."""
    h = a / n
    x = -a / 2 + h * np.arange(n)
    box = (np.abs(x) < 0.48 * a).astype(np.float64)
    z = x[p0:p1]
    layer = np.clip(((z / a + 0.48) / 0.96 * len(values)).astype(int), 0, len(values) - 1)
    vz = np.asarray(values, dtype=np.float64)[layer] * (np.abs(z) < 0.48 * a)
    nu = box[:, None, None] * box[None, :, None] * vz[None, None, :]
    return np.ascontiguousarray(nu.reshape(-1, order="F"))
