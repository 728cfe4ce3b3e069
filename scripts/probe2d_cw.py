"""2-D apply with the P2 -> P3 intermediate interleaved over cw adjacent x-slots (LS_C_INTERLEAVE): timing + identical bits."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fast_solver_lippmann_schwinger_b200 as ls
from fast_solver_lippmann_schwinger_b200.problems import gv_problem_2d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cws = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
nu, gfft, k, h = gv_problem_2d(n)
rng = np.random.default_rng(1234)
b = rng.standard_normal(n * n) + 1j * rng.standard_normal(n * n)
db = ls.DeviceBuffer.from_host(b); dy = ls.DeviceBuffer(b.nbytes)
for cw in cws:
    os.environ["LS_C_INTERLEAVE"] = str(cw)
    M = ls.FastM(gfft, nu, 4 * n, 4 * n, n, n, k, quadRule="Greengard_Vico")
    for _ in range(3):
        M.mul_(dy, db)
    M.sync()
    digest = hashlib.sha1(dy.to_host().tobytes()).hexdigest()[:12]
    reps = 30
    M.profile_enable(True)
    M.timer_start()
    for _ in range(reps):
        M.mul_(dy, db)
    ms = M.timer_stop() / reps
    ph, cnt = M.profile_read(3)
    print("n=%d cw=%d apply %.4f ms -> %.1f applies/s  phases %s  sha1 %s" % (
        n, cw, ms, 1e3 / ms, ["%.4f" % (p / reps) for p in ph], digest), flush=True)
    M.destroy()
