"""Multi-rank tests: world_size-2 gloo run of the host-side sharding logic on CPU, and the
NCCL path (sharded 3-D apply + sharded GMRES vs the oracle) when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def _torchrun(nproc, args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + args
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", "2")
    env.setdefault("LS_ORACLE_WORKERS", str(max(1, (os.cpu_count() or 1) // nproc)))    # every rank evaluates the oracle
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_sharding_host_logic_gloo_world2(built_lib):
    p = _torchrun(2, ["cpu", "64"], 29611)
    assert p.returncode == 0 and "DIST_CPU_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


def test_slab_helpers():
    from fast_solver_lippmann_schwinger_b200 import dist as lsd
    assert lsd.slab_range(256, 3, 8) == (96, 128)
    assert lsd.vector_range(4, 4, 8, 1, 2) == (64, 128)
    with pytest.raises(ValueError):
        lsd.slab_range(100, 0, 8)
    # all-to-all volume formula of SURVEY.md section 8(d): 256^3 on 8 GPUs -> 0.117 GB per GPU
    assert abs(lsd.exchange_bytes_per_rank(256, 256, 256, 8) / 1e9 - 0.1174) < 1e-3


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


def test_row_slab_blocks_of_the_sparsifier():
    """Host logic of the sharded SpMV: every rank's windowed row block, applied to [halo | slab | halo], gives its
    slab of A @ x; the 27-point structure needs exactly n*m + n + 1 entries from each z-neighbour."""
    from util_sparse import stencil27
    from fast_solver_lippmann_schwinger_b200 import dist as lsd
    n, m, l, P = 6, 5, 8, 4
    A = stencil27(n, m, l, seed=3)
    assert lsd.matrix_halo(A) == n * m + n + 1
    N = n * m * l
    rng = np.random.default_rng(5)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    y = A @ x
    H = lsd.matrix_halo(A)
    for r in range(P):
        a, b = lsd.vector_range(n, m, l, r, P)
        assert H <= b - a
        blk = lsd.local_block_csc(A, a, b, H)
        assert blk.shape == (b - a, b - a + 2 * H)
        xext = np.zeros(b - a + 2 * H, complex)
        lo, hi = max(a - H, 0), min(b + H, N)
        xext[lo - (a - H): lo - (a - H) + hi - lo] = x[lo:hi]
        assert np.allclose(blk @ xext, y[a:b], rtol=0, atol=1e-13)
    with pytest.raises(ValueError):
        lsd.local_block_csc(A, 0, n * m, 3)          # halo too small for the stencil


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_apply_and_gmres_nccl(world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    p = _torchrun(world, ["gpu", "64"], 29620 + world, timeout=900)
    assert p.returncode == 0 and "DIST_GPU_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
