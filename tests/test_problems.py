"""The product-side input builders (bench / examples) against the oracle's restatement of the reference's setup."""
import numpy as np

from oracle import ls_oracle as O


def test_gv_spectrum_2d_quadrant_mirror_is_bit_identical():
    from fast_solver_lippmann_schwinger_b200.problems import gv_problem_2d, gv_spectrum_2d
    for n, m in ((16, 16), (12, 20)):
        h = 1.0 / n
        k = 2 * np.pi / (9.3 * h)
        G = gv_spectrum_2d(n, m, h, k)
        Go = O.gv_spectrum_2d(n, m, h, k)
        assert G.shape == (4 * n, 4 * m) and np.array_equal(G, Go)
    nu, G, k, h = gv_problem_2d(32)
    x, ho, ko, Mo = O.pow2_problem_2d(32)
    assert h == ho and k == ko and np.array_equal(nu, Mo.nu) and np.array_equal(G, Mo.GFFT)


def test_nu_gaussian_3d_grid():
    from fast_solver_lippmann_schwinger_b200.problems import nu_gaussian_3d_grid
    n = 12
    h = 1.0 / n
    x = -0.5 + h * np.arange(n)
    X, Y, Z = O.grid3d(x, x, x)
    assert np.allclose(nu_gaussian_3d_grid(n), O.nu_gaussian_3d(X, Y, Z), rtol=1e-14, atol=1e-18)


def test_nu_plasma_2d_matches_the_oracle():
    """Config 3 input builder (tests/plasma_example.jl:53-68) against the oracle's restatement."""
    from oracle import ls_oracle as O
    from fast_solver_lippmann_schwinger_b200.problems import nu_plasma_2d
    n = 96
    x = -0.5 + np.arange(n) / n
    X, Y = O.grid2d(x, x)
    a, b = nu_plasma_2d(X, Y), O.nu_plasma_2d(X, Y)
    assert np.abs(a - b).max() <= 1e-15 * np.abs(b).max()
    assert (a == 0).sum() > 100 and np.unique(a).size > 100


def test_nu_layered_3d_slab_is_a_slab_of_the_full_grid():
    from fast_solver_lippmann_schwinger_b200.problems import nu_layered_3d_slab
    n = 16
    full = nu_layered_3d_slab(n, 0, n)
    assert full.shape == (n ** 3,) and set(np.unique(full)) == {0.0, 0.02, 0.05, 0.08, 0.10}
    parts = [nu_layered_3d_slab(n, p0, p0 + 4) for p0 in range(0, n, 4)]
    assert np.array_equal(np.concatenate(parts), full)
    cube = full.reshape(n, n, n, order="F")
    assert np.all(cube[0] == 0) and np.all(cube[:, 0] == 0) and np.all(cube[:, :, 0] == 0)       # zero outside the box
    assert all(np.unique(cube[1:, 1:, p]).size == 1 for p in range(1, n))                          # piecewise constant in z
