"""GPU parity of the 2-D operator apply against the CPU oracle (through the C ABI).

Tolerance: relative L2 error <= 1e-12 per apply (BASELINE.json north_star).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _problem(n, m=None):
    from oracle import ls_oracle as O
    m = n if m is None else m
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    y = -0.5 * m / n + h * np.arange(m)
    k = 2 * np.pi / (10 * h)
    return O.buildFastConvolution(x, y, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024])
def test_fastconvolution_matches_oracle(n):
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    Mo = _problem(n)
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega, quadRule="Greengard_Vico")
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    y_ref = O.fastconvolution(Mo, b)
    y = Mg * b
    assert _rel(y, y_ref) <= TOL
    # the identity part must not hide the convolution: compare the convolution alone
    assert _rel(y - b, y_ref - b) <= 1e-11
    # plane wave (example.jl:76)
    X, _ = O.grid2d(-0.5 + np.arange(n) / n, -0.5 + np.arange(n) / n)
    u = np.exp(1j * Mo.omega * X)
    assert _rel(Mg * u, O.fastconvolution(Mo, u)) <= TOL
    # mul! into a preallocated vector, and FFTconvolution (Q2: no nu, no omega^2)
    Y = np.zeros(n * n, dtype=np.complex128)
    Mg.mul_(Y, b)
    assert np.array_equal(Y, y)
    assert _rel(ls.FFTconvolution(Mg, b), O.FFTconvolution(Mo, b)) <= TOL


@pytest.mark.parametrize("n,m", [(64, 256), (256, 128), (512, 64)])
def test_rectangular_grid(n, m):
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    Mo = _problem(n, m)
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega, quadRule="Greengard_Vico")
    rng = np.random.default_rng(7)
    b = rng.standard_normal(n * m) + 1j * rng.standard_normal(n * m)
    assert _rel(Mg * b, O.fastconvolution(Mo, b)) <= TOL
    with pytest.raises(ls.LSCudaError):
        ls.FFTconvolution(Mg, b)          # Q3: square grids only


def test_linearity_and_device_buffers():
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 256
    Mo = _problem(n)
    Mg = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.n, Mo.m, Mo.omega, quadRule="Greengard_Vico")
    rng = np.random.default_rng(3)
    a = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    al = 0.3 - 1.7j
    lhs = Mg * (a + al * b)
    rhs = Mg * a + al * (Mg * b)
    assert _rel(lhs, rhs) <= 1e-13
    db = ls.DeviceBuffer.from_host(a)
    dy = ls.DeviceBuffer(a.nbytes)
    Mg.mul_(dy, db)
    Mg.sync()
    assert np.array_equal(dy.to_host(), Mg * a)
    # in place (y aliases b)
    Mg.mul_(db, db)
    Mg.sync()
    assert np.array_equal(db.to_host(), Mg * a)


def test_unsupported_and_invalid_inputs():
    import fast_solver_lippmann_schwinger_b200 as ls
    with pytest.raises(ls.LSUnsupported):                               # padded line too long for the general path
        ls.FastM(np.zeros((4000, 8), complex), np.zeros(1000 * 2), 4000, 8, 1000, 2, 1.0, quadRule="Greengard_Vico")
    M = ls.FastM(np.zeros((256, 256), complex), np.zeros(64 * 64), 256, 256, 64, 64, 1.0, quadRule="Greengard_Vico")
    with pytest.raises(ValueError):
        M * np.zeros(5, complex)                                       # DimensionMismatch
    with pytest.raises(TypeError):
        M * np.zeros(64 * 64)                                          # MethodError upstream: not Complex{Float64}
    with pytest.raises(ls.LSCudaError):
        M._apply(np.zeros(64 * 64, complex), None, 7)                  # unknown mode
    with pytest.raises(ValueError):
        ls.FastM(np.zeros((256, 255), complex), np.zeros(64 * 64), 256, 256, 64, 64, 1.0, quadRule="Greengard_Vico")


@pytest.mark.parametrize("n,m", [(128, 128), (64, 256), (512, 512)])
def test_compact_padding_equals_literal_4x(n, m):
    """Default handles restrict the kernel to the lags the cropped apply touches and run with 2x padding;
    pad4=True evaluates the reference's literal 4x zero padding.  Same operator to rounding."""
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    Mo = _problem(n, m)
    A2 = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, m, Mo.omega, quadRule="Greengard_Vico")
    A4 = ls.FastM(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, n, m, Mo.omega, quadRule="Greengard_Vico", pad4=True)
    rng = np.random.default_rng(n + m)
    b = rng.standard_normal(n * m) + 1j * rng.standard_normal(n * m)
    y2, y4 = A2 * b, A4 * b
    ref = O.fastconvolution(Mo, b)
    assert _rel(y2, ref) <= TOL and _rel(y4, ref) <= TOL
    assert _rel(y2 - b, y4 - b) <= 1e-12          # the convolution part alone
    # a spectrum that is NOT mirror symmetric (random): the restriction argument does not need symmetry
    G = (rng.standard_normal((4 * n, 4 * m)) + 1j * rng.standard_normal((4 * n, 4 * m))) / (n * m)
    nu = rng.standard_normal(n * m)
    B2 = ls.FastM(G, nu, 4 * n, 4 * m, n, m, 1.3, quadRule="Greengard_Vico")
    Bo = O.FastM(G, nu, 4 * n, 4 * m, n, m, 1.3, quadRule="Greengard_Vico")
    assert _rel(B2 * b, O.fastconvolution(Bo, b)) <= TOL


@pytest.mark.parametrize("n,m", [(64, 64), (256, 128), (100, 100)])
def test_spectrum_generated_on_the_device(n, m):
    """ls_op2d_create_gv: Gtruncated2D (Functions.jl:40-42) evaluated on the device (own J0 / J1 expansions, host Hankel
    scalars) gives the same operator as the host-built GFFT, on the power-of-two and on the general-size path."""
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    y = -0.5 * m / n + h * np.arange(m)
    k = 2 * np.pi / (8.3 * h)          # 10 points per wavelength puts grid frequencies exactly on s == k at n = 100 (Q7: NaN upstream)
    Mo = O.buildFastConvolution(x, y, h, k, O.nu_gaussian_2d, quadRule="Greengard_Vico")
    # L, Lp exactly as buildFastConvolution takes them (FastConvolution.jl:187-188)
    Lp = 4.0 * (x[-1] - x[0] + h)
    L = 1.5 * (x[-1] - x[0] + h)
    Mg = ls.FastM(None, Mo.nu, 4 * n, 4 * m, n, m, k, quadRule="Greengard_Vico", L=L, Lp=Lp)
    rng = np.random.default_rng(n + m)
    b = rng.standard_normal(n * m) + 1j * rng.standard_normal(n * m)
    assert _rel(Mg * b, O.fastconvolution(Mo, b)) < 1e-12
    if n == m:
        assert _rel(ls.FFTconvolution(Mg, b), O.FFTconvolution(Mo, b)) < 1e-12
    Mh = ls.FastM(Mo.GFFT, Mo.nu, 4 * n, 4 * m, n, m, k, quadRule="Greengard_Vico")
    assert _rel(Mg * b, Mh * b) < 1e-13
    with pytest.raises(ValueError):
        ls.FastM(None, Mo.nu, 4 * n, 4 * m, n, m, k, quadRule="Greengard_Vico")          # needs L, Lp


def test_q7_grid_frequency_on_the_wave_number_is_guarded():
    """Q7: Gtruncated2D divides by s^2 - k^2; at n = 100 with 10 points per wavelength the lattice points (40, 0), (24, 32), ...
    sit exactly on s == k and the reference's spectrum holds NaN there.  The device generator moves such a frequency by
    sqrt(eps) (removable singularity), so the operator stays finite - and equals the host-built one wherever that is finite."""
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200.problems import gv_spectrum_2d, nu_gaussian_2d
    n = 100
    h = 1.0 / n
    k = 2 * np.pi / (10 * h)
    G = gv_spectrum_2d(n, n, h, k)
    assert not np.isfinite(G).all()
    x = -0.5 + h * np.arange(n)
    X = np.repeat(x[:, None], n, axis=1).reshape(-1, order="F")
    Y = np.repeat(x[None, :], n, axis=0).reshape(-1, order="F")
    M = ls.FastM(None, nu_gaussian_2d(X, Y), 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico", L=1.5 * n * h, Lp=4.0 * n * h)
    b = np.random.default_rng(0).standard_normal(n * n) + 0j
    assert np.isfinite(M * b).all()
