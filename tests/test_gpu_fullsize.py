"""Full-size (BASELINE.json sizes) checks through size-independent properties.

The oracle's literal apply is too slow / too large at 2048^2 and 256^3 for a test, so these use
  * an independent O(N)-per-point direct sum against the spatial kernel g = ifft2(GFFT) (2-D),
  * reciprocity  (M d_a - d_a)[b] / nu_a == (M d_b - d_b)[a] / nu_b   (G is symmetric),
  * linearity,
  * agreement of the two independent device code paths (pruned pow2 vs Bluestein) at n = 512.
Tolerance 1e-12 relative (1e-10 where a single output entry is compared)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("n", [2048, 4096])          # headline size and the largest size the fast path serves
def test_2d_sample_points_against_direct_sum(n):
    import scipy.fft as sfft
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200.problems import gv_problem_2d
    nu, gfft, k, h = gv_problem_2d(n)
    M = ls.FastM(gfft, nu, 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico")
    # spatial kernel of the padded circular convolution: what fft -> .*GFFT (shifted) -> ifft applies
    g = sfft.ifft2(sfft.ifftshift(gfft), workers=-1)
    del gfft
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    y = M * b
    f = (nu * b).reshape((n, n), order="F")
    pts = [(0, 0), (n - 1, n - 1), (n // 2 - 1, n // 2), (517, n - 49), (n - 1, 3), (n // 2, n // 2)]
    ne = 4 * n
    ii = np.arange(n)
    for (i, j) in pts:
        gi = g[(i - ii) % ne][:, (j - ii) % ne]              # g[i - i', j - j']
        ref = b[i + n * j] + k ** 2 * np.sum(gi * f)
        assert abs(y[i + n * j] - ref) <= 1e-10 * abs(ref), (i, j)
    # linearity at full size
    c = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    al = -0.4 + 1.1j
    assert _rel(M * (b + al * c), y + al * (M * c)) <= 1e-13
    # reciprocity
    a_idx, b_idx = (n // 2 - 24) + n * (n // 2 - 14), (n // 2 + 16) + n * (n // 2 - 34)
    ea = np.zeros(n * n, complex); ea[a_idx] = 1.0
    eb = np.zeros(n * n, complex); eb[b_idx] = 1.0
    ra = (M * ea)[b_idx] / nu[a_idx]
    rb = (M * eb)[a_idx] / nu[b_idx]
    assert abs(ra - rb) <= 1e-11 * abs(ra)


def test_2d_fast_and_general_paths_agree_at_512():
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200.problems import gv_problem_2d
    n = 512
    nu, gfft, k, h = gv_problem_2d(n)
    fast = ls.FastM(gfft, nu, 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico")
    gen = ls.FastM(gfft, nu, 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico", force_generic=True)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    assert _rel(gen * b, fast * b) <= 1e-12


@pytest.mark.parametrize("n", [256])
def test_3d_full_size_properties(n):
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200.problems import nu_gaussian_3d_grid
    h = 1.0 / n
    k = 2 * np.pi / (10 * h)
    nu = nu_gaussian_3d_grid(n)
    M = ls.FastM3D(None, nu, 4 * n, 4 * n, 4 * n, n, n, n, k, L=1.8 * n * h, Lp=4.0 * n * h)
    N = n ** 3
    rng = np.random.default_rng(8)
    b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    c = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    al = 0.9 - 0.3j
    yb = M * b
    assert np.all(np.isfinite(yb.view(np.float64)))
    assert _rel(M * (b + al * c), yb + al * (M * c)) <= 1e-13
    mid = n // 2
    a_idx = mid + n * (mid + 3) + n * n * (mid - 5)
    b_idx = (mid - 9) + n * (mid + 1) + n * n * (mid + 7)
    ea = np.zeros(N, complex); ea[a_idx] = 1.0
    eb = np.zeros(N, complex); eb[b_idx] = 1.0
    ra = (M * ea)[b_idx] / nu[a_idx]
    rb = (M * eb)[a_idx] / nu[b_idx]
    assert abs(ra - rb) <= 1e-11 * abs(ra)
    # the impulse response is the free-space Green's function exp(ikr)/(4 pi r) h^3 (quadrature accuracy)
    x = -0.5 + h * np.arange(n)
    pa = np.array([x[mid], x[mid + 3], x[mid - 5]]); pb = np.array([x[mid - 9], x[mid + 1], x[mid + 7]])
    r = np.linalg.norm(pa - pb)
    green = np.exp(1j * k * r) / (4 * np.pi * r) * h ** 3
    assert abs(ra / k ** 2 - green) <= 2e-2 * abs(green)
