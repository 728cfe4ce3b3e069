"""GPU parity of the 3-D operator apply against the CPU oracle (rel. L2 <= 1e-12 per apply)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _problem(n, l=None, ppw=10.0):
    from oracle import ls_oracle as O
    l = n if l is None else l
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    z = -0.5 * l / n + h * np.arange(l)
    k = 2 * np.pi / (ppw * h)
    return h, k, O.buildFastConvolution3D(x, x, z, h, k, O.nu_gaussian_3d)


@pytest.mark.parametrize("n,l", [(64, 64), (64, 128), (128, 64)])
def test_mul_and_fftconvolution_match_oracle(n, l):
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    h, k, Mo = _problem(n, l)
    Mg = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, Mo.n, Mo.m, Mo.l, Mo.omega)
    N = n * n * l
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    y_ref = Mo * b
    y = Mg * b
    assert _rel(y, y_ref) <= TOL
    assert _rel(y - b, y_ref - b) <= 1e-11
    assert _rel(ls.FFTconvolution(Mg, b), O.FFTconvolution3D(Mo, b)) <= TOL
    # device-generated Greengard-Vico spectrum (what the 256^3 / 512^3 configs use)
    Mgen = ls.FastM3D(None, Mo.nu, Mo.ne, Mo.me, Mo.le, Mo.n, Mo.m, Mo.l, Mo.omega, L=1.8 * n * h, Lp=4.0 * n * h)
    assert _rel(Mgen * b, y_ref) <= TOL
    assert Mg.size(1) == N and Mg.eltype() == np.complex128


def test_3d_device_buffers_inplace_and_errors():
    import fast_solver_lippmann_schwinger_b200 as ls
    n = 64
    h, k, Mo = _problem(n)
    Mg = ls.FastM3D(None, Mo.nu, 4 * n, 4 * n, 4 * n, n, n, n, k, L=1.8 * n * h, Lp=4.0 * n * h)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(n ** 3) + 1j * rng.standard_normal(n ** 3)
    y = Mg * b
    db = ls.DeviceBuffer.from_host(b)
    Mg.mul_(db, db)
    Mg.sync()
    assert np.array_equal(db.to_host(), y)
    with pytest.raises(ls.LSCudaError):          # n != m: the reference pads (ne, ne, le)
        ls.FastM3D(None, np.zeros(64 * 128 * 64), 256, 512, 256, 64, 128, 64, 1.0, L=1.0, Lp=4.0)
    with pytest.raises(ls.LSUnsupported):
        ls.FastM3D(None, np.zeros(900 * 900 * 2), 3600, 3600, 8, 900, 900, 2, 1.0, L=1.8, Lp=4.0)   # too long for the general path


def test_long_z_lines_512():
    """k_mid_fused<512, mode B> (the 512^3 / 8-GPU configuration's z pass) against the oracle."""
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    n, l = 64, 512
    h = 1.0 / l
    x = -0.5 * n / l + h * np.arange(n)
    z = -0.5 + h * np.arange(l)
    k = 2 * np.pi / (10 * h)
    nu = lambda X, Y, Z: 0.3 * np.exp(-40 * (16 * X ** 2 + 16 * Y ** 2 + Z ** 2))
    Mo = O.buildFastConvolution3D(x, x, z, h, k, nu)
    Mg = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k)
    rng = np.random.default_rng(99)
    b = rng.standard_normal(n * n * l) + 1j * rng.standard_normal(n * n * l)
    assert _rel(Mg * b, Mo * b) <= TOL


def test_y_lines_512_properties():
    """k_fwd_pruned / k_inv_pruned<512, mode B> on a 512 x 512 x 64 grid: too large for the oracle,
    checked by reciprocity and linearity."""
    import fast_solver_lippmann_schwinger_b200 as ls
    n, l = 512, 64
    h = 1.0 / n
    k = 2 * np.pi / (10 * h)
    x = -0.5 + h * np.arange(n)
    z = -0.5 * l / n + h * np.arange(l)
    g = np.exp(-40 * x ** 2)
    gz = np.exp(-40 * (8 * z) ** 2)
    nu = np.ascontiguousarray((0.3 * g[:, None, None] * g[None, :, None] * gz[None, None, :]).reshape(-1, order="F"))
    M = ls.FastM3D(None, nu, 4 * n, 4 * n, 4 * l, n, n, l, k, L=1.8 * n * h, Lp=4.0 * n * h)
    N = n * n * l
    rng = np.random.default_rng(3)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    c = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    assert _rel(M * (b + 2j * c), M * b + 2j * (M * c)) <= 1e-13
    ia, ib = (250, 260, 30), (262, 247, 35)
    a_idx = ia[0] + n * ia[1] + n * n * ia[2]
    b_idx = ib[0] + n * ib[1] + n * n * ib[2]
    ea = np.zeros(N, complex); ea[a_idx] = 1.0
    eb = np.zeros(N, complex); eb[b_idx] = 1.0
    ra = (M * ea)[b_idx] / nu[a_idx]
    rb = (M * eb)[a_idx] / nu[b_idx]
    assert abs(ra - rb) <= 1e-11 * abs(ra)
    # (no Green's-function check here: upstream uses the x extent for every wave-number axis
    #  (FastConvolution3D.jl:72-79), so for l != n the kernel is not the free-space one - a reference
    #  quirk both the oracle and the device generator reproduce; the cubic case is checked in
    #  tests/test_gpu_fullsize.py)


@pytest.mark.parametrize("n,l", [(48, 48), (20, 36)])
def test_general_sizes_example3d_as_shipped(n, l):
    """examples/example3D.jl:20-54 ships n = 48 (h = 1/48, k = 1/h, padded 192^3): the general-size path
    (Bluestein lines), with the host spectrum and with the device-generated one."""
    from oracle import ls_oracle as O
    import fast_solver_lippmann_schwinger_b200 as ls
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    z = -0.5 * l / n + h * np.arange(l)
    k = 1.0 / h
    Mo = O.buildFastConvolution3D(x, x, z, h, k, O.nu_gaussian_3d)
    Mg = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k)
    N = n * n * l
    rng = np.random.default_rng(48)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    y_ref = Mo * b
    assert _rel(Mg * b, y_ref) <= TOL
    assert _rel(ls.FFTconvolution(Mg, b), O.FFTconvolution3D(Mo, b)) <= TOL
    Mgen = ls.FastM3D(None, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k, L=1.8 * n * h, Lp=4.0 * n * h)
    assert _rel(Mgen * b, y_ref) <= TOL
    if n == 48:
        # rhs and the unpreconditioned GMRES of example3D.jl:71-78 against the oracle history
        from oracle.gmres_is import gmres as gmres_oracle
        X, Y, Z = O.grid3d(x, x, z)
        u_inc = np.exp(1j * k * X)
        rhs = -(Mo * u_inc - u_inc)
        xo, hist_o, _, _ = gmres_oracle(np.zeros(N, complex), lambda v: Mo * v, rhs, maxiter=25)
        xg, hg = ls.gmres_(np.zeros(N, complex), Mg, rhs, maxiter=25, log=True)
        assert hg.iters == len(hist_o)
        assert np.max(np.abs(hg["resnorm"] - hist_o) / hist_o) < 1e-8


@pytest.mark.parametrize("n,l", [(64, 64), (64, 128)])
def test_compact_padding_equals_literal_4x_3d(n, l):
    """Default 3-D handles run with 2x padding on the kernel restricted to the needed lags; pad4=True is the
    reference's literal 4x evaluation."""
    import fast_solver_lippmann_schwinger_b200 as ls
    h, k, Mo = _problem(n, l)
    A2 = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k)
    A4 = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k, pad4=True)
    N = n * n * l
    rng = np.random.default_rng(n + l)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    y2, y4, ref = A2 * b, A4 * b, Mo * b
    assert _rel(y2, ref) <= TOL and _rel(y4, ref) <= TOL
    assert _rel(y2 - b, y4 - b) <= 1e-12
    assert A2.launch_count() == 5 and A4.launch_count() == 5


@pytest.mark.parametrize("chunks", [2, 8])
@pytest.mark.parametrize("pad4", [False, True])
def test_x_slot_chunks_give_the_same_apply(chunks, pad4, monkeypatch):
    """The sharded operator pipelines its transposes over x-slot chunks (chunk-major exchange buffers and
    spectrum, P2-P4 per chunk).  The same code path runs on one GPU with LS_OP3D_CHUNKS: bit-identical."""
    import fast_solver_lippmann_schwinger_b200 as ls
    n, l = 64, 128
    h, k, Mo = _problem(n, l)
    N = n * n * l
    rng = np.random.default_rng(77)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    ref = Mo * b
    monkeypatch.delenv("LS_OP3D_CHUNKS", raising=False)
    A1 = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k, pad4=pad4)
    y1 = A1 * b
    monkeypatch.setenv("LS_OP3D_CHUNKS", str(chunks))
    Ac = ls.FastM3D(Mo.GFFT, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k, pad4=pad4)
    Ag = ls.FastM3D(None, Mo.nu, Mo.ne, Mo.me, Mo.le, n, n, l, k, L=1.8 * n * h, Lp=4.0 * n * h, pad4=pad4)
    yc = Ac * b
    assert Ac.launch_count() == 2 + 3 * chunks
    assert _rel(yc, ref) <= TOL and _rel(Ag * b, ref) <= TOL
    assert np.array_equal(yc, y1)
    assert _rel(ls.FFTconvolution(Ac, b), ls.FFTconvolution(A1, b)) == 0.0
