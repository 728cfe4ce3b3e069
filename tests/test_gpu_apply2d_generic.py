"""GPU parity of the general-size 2-D path (Bluestein lines): the sizes the reference ships
(examples/example.jl n = 201, Greengard_Vico padded 804) and the trapezoidal rule.
Tolerance: relative L2 <= 1e-12 per apply."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _grid(n, m, h):
    x = -0.5 * (n - 1) * h + h * np.arange(n)
    y = -0.5 * (m - 1) * h + h * np.arange(m)
    return x, y


def test_example_jl_as_shipped():
    """examples/example.jl:30-54,76-77: h = 0.005, n = 201, k = 1/h, Greengard_Vico (ne = 804)."""
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    x, Mo = O.example_problem_2d(h=0.005)
    assert Mo.n == 201 and Mo.ne == 804
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega, quadRule="Greengard_Vico")
    X, Y = O.grid2d(x, x)
    u_inc = np.exp(1j * Mo.omega * X)
    assert _rel(Mg * u_inc, O.fastconvolution(Mo, u_inc)) <= TOL
    rhs_ref = -Mo.omega ** 2 * O.FFTconvolution(Mo, Mo.nu * u_inc)
    rhs = -Mo.omega ** 2 * ls.FFTconvolution(Mg, Mo.nu * u_inc)
    assert _rel(rhs, rhs_ref) <= TOL
    rng = np.random.default_rng(1)
    b = rng.standard_normal(201 * 201) + 1j * rng.standard_normal(201 * 201)
    assert _rel(Mg * b, O.fastconvolution(Mo, b)) <= TOL
    # unpreconditioned GMRES of example.jl:91 - history against the oracle
    from oracle.gmres_is import gmres as gmres_oracle
    xo = np.zeros(201 * 201, complex)
    xo, hist_o, conv_o, mv_o = gmres_oracle(xo, lambda v: O.fastconvolution(Mo, v), rhs_ref, maxiter=30)
    xg = np.zeros(201 * 201, complex)
    xg, hg = ls.gmres_(xg, Mg, rhs_ref, maxiter=30, log=True)
    assert hg.iters == len(hist_o)
    assert np.max(np.abs(hg["resnorm"] - hist_o) / hist_o) < 1e-8


@pytest.mark.parametrize("n,m", [(21, 21), (101, 101), (51, 77)])
def test_trapezoidal_rule(n, m):
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    h = 1.0 / (max(n, m) - 1)
    x, y = _grid(n, m, h)
    k = 1.0 / h
    Mo = O.buildFastConvolution(x, y, h, k, O.nu_gaussian_2d, quadRule="trapezoidal")
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega)       # ctor default quadRule = "trapezoidal"
    rng = np.random.default_rng(n)
    b = rng.standard_normal(n * m) + 1j * rng.standard_normal(n * m)
    assert _rel(Mg * b, O.fastconvolution(Mo, b)) <= TOL
    if n == m:
        assert _rel(ls.FFTconvolution(Mg, b), O.FFTconvolution(Mo, b)) <= TOL      # Q2: nu applied in this branch
    if n == 21:
        # independent of any FFT: the dense Green matrix of buildConvMatrix (FastConvolution.jl:497-513)
        X, Y = O.grid2d(x, y)
        G = O.buildConvMatrix(k, X, Y, 1 - 0.892j, h)
        assert _rel(Mg * b, b + k ** 2 * (G @ (Mo.nu * b))) <= TOL


@pytest.mark.parametrize("n,m", [(100, 37), (96, 96), (5, 9), (640, 48)])
def test_greengard_vico_general_sizes(n, m):
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    h = 1.0 / max(n, m)
    x, y = _grid(n, m, h)
    k = 2 * np.pi / (8.3 * h)      # 8.0 would put a grid frequency exactly on |kappa| = k: NaN upstream too (Q7)
    Mo = O.buildFastConvolution(x, y, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega, quadRule="Greengard_Vico")
    rng = np.random.default_rng(n * m)
    b = rng.standard_normal(n * m) + 1j * rng.standard_normal(n * m)
    assert _rel(Mg * b, O.fastconvolution(Mo, b)) <= TOL


def test_general_path_agrees_with_fast_path():
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 128
    x, h, k, Mo = O.pow2_problem_2d(n)
    fast = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, n, k, quadRule="Greengard_Vico")
    gen = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, n, k, quadRule="Greengard_Vico", force_generic=True)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    assert _rel(gen * b, fast * b) <= 1e-13
    assert gen.launch_count() == 3 and fast.launch_count() == 3
