"""2-D apply: variants of the fused middle pass (LS_P2_KERNEL / LS_P2_GLOAD / LS_P2_MINB), one subprocess per variant
(the choices are read once per process).  Prints per-phase times and a digest of the output (bits must agree)."""
import hashlib
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(n):
    import numpy as np
    import fast_solver_lippmann_schwinger_b200 as ls
    from fast_solver_lippmann_schwinger_b200.problems import gv_problem_2d
    nu, gfft, k, h = gv_problem_2d(n)
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
    db = ls.DeviceBuffer.from_host(b); dy = ls.DeviceBuffer(b.nbytes)
    M = ls.FastM(gfft, nu, 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico")
    for _ in range(5):
        M.mul_(dy, db)
    M.sync()
    digest = hashlib.sha1(dy.to_host().tobytes()).hexdigest()[:12]
    reps = 40
    M.profile_enable(True)
    M.timer_start()
    for _ in range(reps):
        M.mul_(dy, db)
    ms = M.timer_stop() / reps
    ph, cnt = M.profile_read(3)
    print("n=%d %s apply %.4f ms -> %.1f applies/s  P1/P2/P3 %s  sha1 %s" % (
        n, os.environ.get("VARIANT", ""), ms, 1e3 / ms, ["%.4f" % (p / reps) for p in ph], digest), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
        sys.exit(0)
    sizes = [int(a) for a in sys.argv[1:]] or [2048]
    variants = [("fused(r1)", {"LS_P2_KERNEL": "0"}),
                ("swap,3cta,regG", {"LS_P2_KERNEL": "1", "LS_P2_GLOAD": "0", "LS_P2_MINB": "3"}),
                ("swap,3cta,L2G", {"LS_P2_KERNEL": "1", "LS_P2_GLOAD": "1", "LS_P2_MINB": "3"}),
                ("swap,2cta,regG", {"LS_P2_KERNEL": "1", "LS_P2_GLOAD": "0", "LS_P2_MINB": "2"}),
                ("swap,2cta,L2G", {"LS_P2_KERNEL": "1", "LS_P2_GLOAD": "1", "LS_P2_MINB": "2"})]
    for n in sizes:
        for name, env in variants:
            e = dict(os.environ); e.update(env); e["VARIANT"] = name
            subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(n)], env=e, check=False)
