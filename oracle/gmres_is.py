"""CPU oracle: IterativeSolvers.jl ``gmres!`` restated.  TEST INFRASTRUCTURE ONLY.

IterativeSolvers.jl is NOT under /root/reference (README.md:8 only names the package; there
is no Project.toml / Manifest.toml, so the version is unpinned).  This file restates the
published algorithm of its ``src/gmres.jl`` / ``src/orthogonalize.jl`` / ``src/hessenberg.jl``
(v0.8 / v0.9 line) as summarised in SURVEY.md Appendix A, anchored on the reference's call
sites:  examples/example.jl:85,91  (``gmres!(u, fastconv, rhs, Pl=precond, log=true)``,
history read at :86), examples/example3D.jl:78, tests/plasma_example.jl:164,176.

PARITY UNPINNED (no Julia in this image, no golden histories upstream).  Pinned instead by
tests/test_oracle.py: the history equals the exact minimal preconditioned residual over the
Krylov space (dense least squares), and x solves the system at convergence.

Semantics kept from upstream:
  * left preconditioning only (Pr = Identity): residuals are ||Pl^-1 (b - A x)||_2;
  * restart = min(20, N), maxiter = N counts INNER iterations, reltol = sqrt(eps),
    tol = max(reltol * beta0, abstol) with beta0 the first preconditioned residual;
  * modified Gram-Schmidt: h_i = dot(V_i, w) (conjugates V_i), w -= h_i V_i;
  * residual estimate through the null-vector recurrence (``update_residual!``);
  * on restart the true preconditioned residual is recomputed (one extra mul! + ldiv!),
    the logged value of that iteration stays the estimate.
"""
from __future__ import annotations

import numpy as np


def _givens(f, g):
    """LinearAlgebra.givensAlgorithm semantics: c real, s complex, [c s; -conj(s) c]*[f;g]=[r;0]."""
    if g == 0:
        return 1.0, 0.0 + 0.0j, f
    if f == 0:
        return 0.0, np.conj(g) / abs(g), abs(g)
    nf = abs(f)
    d = np.hypot(nf, abs(g))
    c = nf / d
    s = (f / nf) * np.conj(g) / d
    r = (f / nf) * d
    return c, s, r


def solve_least_squares(H, beta, k):
    """hessenberg.jl ``ldiv!(::FastHessenberg, rhs)`` on H[1:k, 1:k-1]; returns y (len k-1)."""
    width = k - 1
    Hh = np.array(H[:k, :width], dtype=np.complex128)
    rhs = np.zeros(k, dtype=np.complex128)
    rhs[0] = beta
    for i in range(width):
        c, s, _ = _givens(Hh[i, i], Hh[i + 1, i])
        Hh[i, i] = c * Hh[i, i] + s * Hh[i + 1, i]
        for j in range(i + 1, width):
            tmp = -np.conj(s) * Hh[i, j] + c * Hh[i + 1, j]
            Hh[i, j] = c * Hh[i, j] + s * Hh[i + 1, j]
            Hh[i + 1, j] = tmp
        tmp = -np.conj(s) * rhs[i] + c * rhs[i + 1]
        rhs[i] = c * rhs[i] + s * rhs[i + 1]
        rhs[i + 1] = tmp
    y = np.zeros(width, dtype=np.complex128)
    for i in range(width - 1, -1, -1):
        y[i] = (rhs[i] - Hh[i, i + 1:width] @ y[i + 1:]) / Hh[i, i]
    return y


def _orthogonalize(V, k, w, orth_meth):
    """orthogonalize.jl ``orthogonalize_and_normalize!``: returns (h[0:k], nrm, w/nrm)."""
    if orth_meth == "ModifiedGramSchmidt":
        h = np.zeros(k, dtype=np.complex128)
        for i in range(k):
            h[i] = np.vdot(V[:, i], w)
            w = w - h[i] * V[:, i]
        nrm = np.linalg.norm(w)
        return h, nrm, w * (1.0 / nrm)
    Vk = V[:, :k]
    h = Vk.conj().T @ w                          # mul!(h, adjoint(V), w)
    w = w - Vk @ h                               # mul!(w, V, h, -1, 1)
    nrm = np.linalg.norm(w)
    if orth_meth == "DGKS":
        eta = 1.0 / np.sqrt(2.0)
        projection_size = np.linalg.norm(h)
        while nrm < eta * projection_size:
            correction = Vk.conj().T @ w
            projection_size = np.linalg.norm(correction)
            w = w - Vk @ correction
            h = h + correction
            nrm = np.linalg.norm(w)
    elif orth_meth != "ClassicalGramSchmidt":
        raise ValueError(orth_meth)
    return h, nrm, w * (1.0 / nrm)


def gmres(x, A_mul, b, Pl_ldiv=None, abstol=0.0, reltol=None, restart=None, maxiter=None,
          initially_zero=False, orth_meth="ModifiedGramSchmidt"):
    """``gmres!(x, A, b; Pl, abstol, reltol, restart, maxiter, log=true)``.

    ``A_mul(v) -> A v``; ``Pl_ldiv(v) -> Pl^-1 v`` (None = Identity).
    Returns (x, history, converged, mv_products); ``x`` is updated in place.
    """
    N = b.shape[0]
    reltol = np.sqrt(np.finfo(np.float64).eps) if reltol is None else reltol
    restart = min(20, N) if restart is None else restart
    maxiter = N if maxiter is None else maxiter
    pl = (lambda v: v) if Pl_ldiv is None else Pl_ldiv

    V = np.zeros((N, restart + 1), dtype=np.complex128, order="F")
    H = np.zeros((restart + 1, restart), dtype=np.complex128)
    nullvec = np.ones(restart + 1, dtype=np.complex128)
    mv = 1 if initially_zero else 0

    def init(first):
        v = np.array(b, dtype=np.complex128)
        if not (first and initially_zero):
            v = v - A_mul(x)
        v = pl(v)
        beta = np.linalg.norm(v)
        V[:, 0] = v * (1.0 / beta)
        return beta

    beta = init(True)
    current = beta
    accumulator = 1.0
    rbeta = beta
    tol = max(reltol * current, abstol)
    k = 1
    history = []
    iteration = 0

    def converged():
        return current <= tol

    # gmres.jl: done(g, it) = it >= maxiter || converged(g)
    while not (iteration >= maxiter or converged()):
        # expand!
        w = pl(A_mul(V[:, k - 1]))
        mv += 1
        # orthogonalize_and_normalize!  (ModifiedGramSchmidt unless asked otherwise)
        hcol, nrm, wn = _orthogonalize(V, k, w, orth_meth)
        H[:k, k - 1] = hcol
        H[k, k - 1] = nrm
        V[:, k] = wn
        # update_residual!
        nullvec[k] = -np.conj(np.vdot(nullvec[:k], H[:k, k - 1]) / H[k, k - 1])
        accumulator += abs(nullvec[k]) ** 2
        current = rbeta / np.sqrt(accumulator)
        k += 1
        # gmres.jl iterate(): x is formed at a restart or at convergence only (if maxiter lands inside a cycle the
        # caller gets the x of the last restart), and the cycle restarts whenever the solve has not converged
        if k == restart + 1 or converged():
            y = solve_least_squares(H, beta, k)
            x += V[:, :k - 1] @ y
            k = 1
            if not converged():
                beta = init(False)
                accumulator = 1.0
                rbeta = beta
                mv += 1
        iteration += 1
        history.append(current)
    return x, np.array(history), current <= tol, mv
