"""Generates the committed golden fixtures from the CPU oracle (the reference itself cannot run
here: Julia absent - see oracle/ls_oracle.py header).  Run:  python tests/golden/make_golden.py

Each .npz holds the inputs needed to rebuild the problem (sizes, k, h, seeds) and the oracle's
outputs, so the GPU tests can check against bytes that do not change when the oracle is edited.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ls_oracle as O          # noqa: E402
from oracle.gmres_is import gmres          # noqa: E402


def apply2d(n, m, seed):
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    y = -0.5 * m / n + h * np.arange(m)
    k = 2 * np.pi / (10 * h)
    M = O.buildFastConvolution(x, y, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    rng = np.random.default_rng(seed)
    b = rng.standard_normal(n * m) + 1j * rng.standard_normal(n * m)
    out = {"n": n, "m": m, "h": h, "k": k, "seed": seed, "y_fastconvolution": O.fastconvolution(M, b)}
    if n == m:
        out["y_FFTconvolution"] = O.FFTconvolution(M, b)
    return out


def apply3d(n, seed):
    x, h, k, M = O.pow2_problem_3d(n)
    rng = np.random.default_rng(seed)
    b = rng.standard_normal(n ** 3) + 1j * rng.standard_normal(n ** 3)
    return {"n": n, "h": h, "k": k, "seed": seed, "y_mul": M * b, "y_FFTconvolution": O.FFTconvolution3D(M, b)}


def gmres2d(n):
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    k = 1.0 / h
    M = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    X, Y = O.grid2d(x, x)
    D0 = O.referenceValsTrapRule()[1][0]
    cache = O.entriesSparseA(k, X, Y, D0, n, n, strict=False)
    As = O.buildSparseA(k, X, Y, D0, n, n, strict=False, _cache=cache)
    AG = O.buildSparseAG(k, X, Y, D0, n, n, strict=False, _cache=cache)
    Msp = (As + k ** 2 * (AG @ sp.diags(M.nu))).tocsc()
    P = O.SparsifyingPreconditioner(Msp, As)
    rhs = -k ** 2 * O.FFTconvolution(M, M.nu * np.exp(1j * k * X))
    x0 = np.zeros(n * n, complex)
    x0, hist_p, conv_p, mv_p = gmres(x0, lambda v: O.fastconvolution(M, v), rhs, Pl_ldiv=P.solve)
    x1 = np.zeros(n * n, complex)
    x1, hist_u, conv_u, mv_u = gmres(x1, lambda v: O.fastconvolution(M, v), rhs)
    cp, rv, nz = O.julia_csc_arrays(As)
    cpm, rvm, nzm = O.julia_csc_arrays(Msp)
    return {"n": n, "h": h, "k": k, "As_colptr": cp, "As_rowval": rv, "As_nzval": nz,
            "Msp_colptr": cpm, "Msp_rowval": rvm, "Msp_nzval": nzm, "rhs": rhs,
            "hist_precond": hist_p, "x_precond": x0, "mv_precond": mv_p,
            "hist_plain": hist_u, "x_plain": x1, "mv_plain": mv_u}


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "apply2d_n64.npz"), **apply2d(64, 64, 1234))
    np.savez_compressed(os.path.join(HERE, "apply2d_n128x64.npz"), **apply2d(128, 64, 77))
    np.savez_compressed(os.path.join(HERE, "apply3d_n64.npz"), **apply3d(64, 4321)) if "--with3d" in sys.argv else None
    np.savez_compressed(os.path.join(HERE, "gmres2d_n64.npz"), **gmres2d(64))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
