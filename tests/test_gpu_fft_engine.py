"""GPU unit test of the register-resident line-FFT engine against numpy's FFT."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [64, 128, 256, 512, 1024, 2048, 4096]


@pytest.mark.parametrize("N", SIZES)
def test_forward_matches_numpy(N):
    from fast_solver_lippmann_schwinger_b200._lib import check, lib, ptr
    rng = np.random.default_rng(N)
    nl = 32
    x = (rng.standard_normal((nl, N)) + 1j * rng.standard_normal((nl, N))).astype(np.complex128)
    out = np.empty_like(x)
    check(lib().ls_test_fft_lines(N, nl, ptr(x), ptr(out), 0))
    ref = np.fft.fft(x, axis=1)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < 1e-14, err


@pytest.mark.parametrize("N", SIZES)
def test_roundtrip(N):
    from fast_solver_lippmann_schwinger_b200._lib import check, lib, ptr
    rng = np.random.default_rng(N + 1)
    nl = 16
    x = (rng.standard_normal((nl, N)) + 1j * rng.standard_normal((nl, N))).astype(np.complex128)
    out = np.empty_like(x)
    check(lib().ls_test_fft_lines(N, nl, ptr(x), ptr(out), 1))
    err = np.linalg.norm(out - x) / np.linalg.norm(x)
    assert err < 1e-14, err
