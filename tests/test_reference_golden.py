"""Parity against vectors produced by the UNMODIFIED reference (julia/gen_golden.jl -> tests/golden/ref_example_*.npy).

The build image has no Julia, so the files are absent in this repository's history and both tests SKIP with that
reason - loudly, because it is exactly what keeps the parity claim at "unpinned" (DESIGN.md section 2).  Once a
maintainer has run `julia julia/gen_golden.jl` the CPU test pins the oracle and the GPU test pins the CUDA path to the
reference itself: one apply and As*b to 1e-12 / 1e-13 relative L2, residual histories to 1e-8 over 50 iterations.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NEEDED = ["meta", "b", "apply", "fftconv", "Asb", "precond_b", "As_colptr", "As_rowval", "As_nzval",
          "Msp_colptr", "Msp_rowval", "Msp_nzval", "rhs", "hist_precond", "hist_plain", "u_precond"]


def _load():
    missing = [n for n in NEEDED if not os.path.exists(os.path.join(GOLD, "ref_example_%s.npy" % n))]
    if missing:
        pytest.skip("REFERENCE-PINNED FIXTURES ABSENT (tests/golden/ref_example_*.npy): the reference is Julia and Julia is "
                    "not installed in the build image; run `julia julia/gen_golden.jl` on a machine that has it. "
                    "Parity stays pinned to the CPU oracle only.")
    return {n: np.load(os.path.join(GOLD, "ref_example_%s.npy" % n)) for n in NEEDED}


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _csc(g, name, N):
    import scipy.sparse as sp
    return sp.csc_matrix((g[name + "_nzval"], g[name + "_rowval"] - 1, g[name + "_colptr"] - 1), shape=(N, N))


def test_oracle_reproduces_the_reference_vectors():
    g = _load()
    from oracle import ls_oracle as O
    from oracle.gmres_is import gmres
    n, m, h, k = int(g["meta"][0]), int(g["meta"][1]), g["meta"][2], g["meta"][3]
    x = -0.5 + h * np.arange(n)
    M = O.buildFastConvolution(x, x, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    assert _rel(O.fastconvolution(M, g["b"]), g["apply"]) < 1e-12
    assert _rel(O.FFTconvolution(M, g["b"]), g["fftconv"]) < 1e-12
    N = n * m
    As, Msp = _csc(g, "As", N), _csc(g, "Msp", N)
    assert _rel(O.csc_matvec(As, g["b"]), g["Asb"]) < 1e-13
    X, Y = O.grid2d(x, x)
    D0 = complex(g["meta"][4], g["meta"][5])
    As_o = O.buildSparseA(k, X, Y, D0, n, m)
    # stencil vectors are singular vectors: equal up to one unit phase per boundary class
    P = O.SparsifyingPreconditioner(Msp, As)
    assert _rel(P.solve(g["b"]), g["precond_b"]) < 1e-9
    Po = O.SparsifyingPreconditioner((As_o + k ** 2 * (O.buildSparseAG(k, X, Y, D0, n, m) @ __import__("scipy.sparse").sparse.diags(M.nu))).tocsc(), As_o)
    assert _rel(Po.solve(g["b"]), g["precond_b"]) < 1e-8
    rhs = -k ** 2 * O.FFTconvolution(M, M.nu * np.exp(1j * k * X))
    assert _rel(rhs, g["rhs"]) < 1e-12
    _, hist, _, _ = gmres(np.zeros(N, complex), lambda v: O.fastconvolution(M, v), g["rhs"], Pl_ldiv=P.solve)
    mm = min(50, len(g["hist_precond"]))
    assert len(hist) == len(g["hist_precond"]) and np.max(np.abs(hist[:mm] - g["hist_precond"][:mm]) / g["hist_precond"][:mm]) < 1e-8
    _, hist0, _, _ = gmres(np.zeros(N, complex), lambda v: O.fastconvolution(M, v), g["rhs"])
    mm = min(50, len(g["hist_plain"]))
    assert len(hist0) == len(g["hist_plain"]) and np.max(np.abs(hist0[:mm] - g["hist_plain"][:mm]) / g["hist_plain"][:mm]) < 1e-8


@pytest.mark.gpu
def test_gpu_path_reproduces_the_reference_vectors():
    g = _load()
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200.problems import gv_spectrum_2d, nu_gaussian_2d
    n, m, h, k = int(g["meta"][0]), int(g["meta"][1]), g["meta"][2], g["meta"][3]
    x = -0.5 + h * np.arange(n)
    X = np.repeat(x[:, None], m, axis=1).reshape(-1, order="F")
    Y = np.repeat(x[None, :], n, axis=0).reshape(-1, order="F")
    M = ls.FastM(gv_spectrum_2d(n, m, h, k), nu_gaussian_2d(X, Y), 4 * n, 4 * m, n, m, k, quadRule="Greengard_Vico")
    assert _rel(M * g["b"], g["apply"]) < 1e-12
    assert _rel(ls.FFTconvolution(M, g["b"]), g["fftconv"]) < 1e-12
    N = n * m
    As, Msp = _csc(g, "As", N), _csc(g, "Msp", N)
    assert _rel(ls.GPUSparseMatrixCSC(As) * g["b"], g["Asb"]) < 1e-13
    for kw in ({"solverType": "GPU", "grid": (n, m)}, {}):
        P = ls.SparsifyingPreconditioner(Msp, As, **kw)
        assert _rel(P.solve(g["b"]), g["precond_b"]) < 1e-9
        u, hist = ls.gmres_(np.zeros(N, complex), M, g["rhs"], Pl=P, log=True)
        mm = min(50, len(g["hist_precond"]))
        assert hist.iters == len(g["hist_precond"])
        assert np.max(np.abs(hist["resnorm"][:mm] - g["hist_precond"][:mm]) / g["hist_precond"][:mm]) < 1e-8
        assert _rel(u, g["u_precond"]) < 1e-7
    _, hist0 = ls.gmres_(np.zeros(N, complex), M, g["rhs"], log=True)
    mm = min(50, len(g["hist_plain"]))
    assert hist0.iters == len(g["hist_plain"])
    assert np.max(np.abs(hist0["resnorm"][:mm] - g["hist_plain"][:mm]) / g["hist_plain"][:mm]) < 1e-8
