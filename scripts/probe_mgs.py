"""MGS sweep and GMRES(20) iteration time with / without the L2 residency hints (LS_MGS_L2HINT), one process per setting."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(n):
    import numpy as np
    import fast_solver_lippmann_schwinger_b200 as ls
    N = n * n
    h = 1.0 / n
    k = 2 * np.pi / (10 * h)
    x = -0.5 + h * np.arange(n)
    X = np.repeat(x[:, None], n, axis=1).reshape(-1, order="F")
    Y = np.repeat(x[None, :], n, axis=0).reshape(-1, order="F")
    from fast_solver_lippmann_schwinger_b200.problems import nu_gaussian_2d
    M = ls.FastM(None, nu_gaussian_2d(X, Y), 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico", L=1.5 * n * h, Lp=4.0 * n * h)
    ws = ls.KrylovWorkspace(N)
    rng = np.random.default_rng(0)
    w = ls.DeviceBuffer.from_host(rng.standard_normal(N) + 1j * rng.standard_normal(N))
    out = []
    for kk in (5, 10, 20):
        V = ls.DeviceBuffer(16 * N * (kk + 1))
        ws.mgs_step(V, N, kk, w)
        ws.timer_start()
        for _ in range(10):
            ws.mgs_step(V, N, kk, w)
        ms = ws.timer_stop() / 10
        out.append("k=%d %.4f ms (%.0f GB/s on (64k+80)N)" % (kk, ms, (64 * kk + 80) * N / ms / 1e6))
        V.free()
    u_inc = np.exp(1j * k * X)
    rhs = -(M * u_inc - u_inc)
    db = ls.DeviceBuffer.from_host(rhs)
    dx = ls.DeviceBuffer.from_host(np.zeros(N, complex))
    ls.gmres_(dx, M, db, reltol=1e-8, maxiter=25, workspace=ws)
    dx = ls.DeviceBuffer.from_host(np.zeros(N, complex))
    t0 = time.perf_counter()
    _, hist = ls.gmres_(dx, M, db, reltol=1e-8, maxiter=1000, log=True, workspace=ws)
    dt = time.perf_counter() - t0
    print("n=%d L2HINT=%s  MGS %s | GMRES %d its %.3f s = %.3f ms/iter last %.3e" % (
        n, os.environ.get("LS_MGS_L2HINT", "0"), "; ".join(out), hist.iters, dt, 1e3 * dt / max(hist.iters, 1), hist["resnorm"][-1]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
        sys.exit(0)
    for n in [int(a) for a in sys.argv[1:]] or [2048]:
        for hint in ("0", "1"):
            e = dict(os.environ); e["LS_MGS_L2HINT"] = hint
            subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(n)], env=e, check=False)
