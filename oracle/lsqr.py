"""Oracle (test infrastructure, CPU): LSQR as called by the reference.

Call sites restated: ``IterativeSolvers.lsqr(LHS, RHS)`` at
``src/QuadraticProgram/QuadraticProgram.jl:488`` and ``lsqr(M, g)`` at
``src/ConicProgram/ConicProgram.jl:323,372``.  IterativeSolvers (compat "0.9",
``Project.toml:20``) is NOT vendored in /root/reference, so this is a
restatement of its published algorithm: Paige & Saunders' LSQR (ACM TOMS 8(1),
1982) with x0 = 0, damp = 0 and the package defaults

    atol = btol = sqrt(eps(Float64)),  conlim = 1/sqrt(eps(Float64)),
    maxiter = max(size(A))

(defaults recalled from IterativeSolvers 0.9 ``lsqr.jl``; parity-unpinned, see
``oracle/__init__.py``).  The GPU path and this oracle are always compared with
the SAME explicit tolerances, as ``BASELINE.json:north_star`` asks ("matched
residual tolerance").

``tests/test_oracle_kat.py`` cross-checks this against
``scipy.sparse.linalg.lsqr`` (an independent transliteration of the same paper)
and against ``numpy.linalg.pinv`` (the minimum-norm limit on singular systems).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)


@dataclass
class LsqrInfo:
    istop: int
    itn: int
    rnorm: float   # estimate of ||b - A x||
    arnorm: float  # estimate of ||A'(b - A x)||
    anorm: float
    acond: float
    xnorm: float


def _as_ops(A):
    if isinstance(A, tuple):  # (matvec, rmatvec, (m, n))
        return A
    m, n = A.shape
    if hasattr(A, "tocsr"):
        Ar = A.tocsr()
        At = A.T.tocsr()
        return (lambda v: Ar @ v), (lambda u: At @ u), (m, n)
    return (lambda v: A @ v), (lambda u: A.T @ u), (m, n)


def lsqr(A, b, atol=SQRT_EPS, btol=SQRT_EPS, conlim=1.0 / SQRT_EPS, maxiter=None,
         return_info=False):
    """Minimise ||A x - b||_2 from x0 = 0 (min-norm limit on singular A).

    ``A`` is a dense ndarray, a scipy sparse matrix, or a triple
    ``(matvec, rmatvec, (m, n))``.  Stop codes follow the paper / the
    reference package: 1 = ||r|| small (test1 <= btol + atol*|A||x|/|b|),
    2 = ||A'r|| small (test2 <= atol), 3 = cond(A) >= conlim, 4/5/6 = the same
    at machine precision, 7 = maxiter.
    """
    matvec, rmatvec, (m, n) = _as_ops(A)
    b = np.asarray(b, dtype=np.float64).ravel()
    if maxiter is None:
        maxiter = max(m, n)
    x = np.zeros(n)
    itn = 0
    istop = 0
    ctol = 1.0 / conlim if conlim > 0 else 0.0
    anorm = acond = ddnorm = res2 = xnorm = xxnorm = z = sn2 = 0.0
    cs2 = -1.0

    # beta*u = b ; alpha*v = A'u     (x0 = 0)
    u = b.copy()
    beta = float(np.linalg.norm(u))
    alpha = 0.0
    v = np.zeros(n)
    if beta > 0:
        u *= 1.0 / beta
        v = rmatvec(u).astype(np.float64, copy=True)
        alpha = float(np.linalg.norm(v))
    if alpha > 0:
        v *= 1.0 / alpha
    w = v.copy()
    arnorm = alpha * beta
    rnorm = beta
    if arnorm == 0:
        info = LsqrInfo(0, 0, rnorm, arnorm, 0.0, 0.0, 0.0)
        return (x, info) if return_info else x
    rhobar = alpha
    phibar = bnorm = beta

    while itn < maxiter:
        itn += 1
        # bidiagonalisation step: beta*u = A v - alpha*u ; alpha*v = A'u - beta*v
        u = matvec(v) - alpha * u
        beta = float(np.linalg.norm(u))
        if beta > 0:
            u *= 1.0 / beta
            anorm = math.sqrt(anorm * anorm + alpha * alpha + beta * beta)
            v = rmatvec(u) - beta * v
            alpha = float(np.linalg.norm(v))
            if alpha > 0:
                v *= 1.0 / alpha
        # damp = 0: first rotation is the identity
        rhobar1 = rhobar
        # plane rotation eliminating the sub-diagonal beta
        rho = math.hypot(rhobar1, beta)
        cs = rhobar1 / rho
        sn = beta / rho
        theta = sn * alpha
        rhobar = -cs * alpha
        phi = cs * phibar
        phibar = sn * phibar
        tau = sn * phi
        # update x, w
        t1 = phi / rho
        t2 = -theta / rho
        dk = w * (1.0 / rho)
        x = x + t1 * w
        w = v + t2 * w
        ddnorm += float(dk @ dk)
        # estimate ||x||
        delta = sn2 * rho
        gambar = -cs2 * rho
        rhs = phi - delta * z
        zbar = rhs / gambar
        xnorm = math.sqrt(xxnorm + zbar * zbar)
        gamma = math.hypot(gambar, theta)
        cs2 = gambar / gamma
        sn2 = theta / gamma
        z = rhs / gamma
        xxnorm += z * z
        # convergence tests
        acond = anorm * math.sqrt(ddnorm)
        rnorm = math.sqrt(phibar * phibar + res2)
        arnorm = alpha * abs(tau)
        test1 = rnorm / bnorm
        test2 = arnorm / (anorm * rnorm) if anorm * rnorm > 0 else math.inf
        test3 = 1.0 / acond if acond > 0 else math.inf
        t1_ = test1 / (1.0 + anorm * xnorm / bnorm)
        rtol = btol + atol * anorm * xnorm / bnorm
        if itn >= maxiter:
            istop = 7
        if 1.0 + test3 <= 1.0:
            istop = 6
        if 1.0 + test2 <= 1.0:
            istop = 5
        if 1.0 + t1_ <= 1.0:
            istop = 4
        if test3 <= ctol:
            istop = 3
        if test2 <= atol:
            istop = 2
        if test1 <= rtol:
            istop = 1
        if istop > 0:
            break
    info = LsqrInfo(istop, itn, rnorm, arnorm, anorm, acond, xnorm)
    return (x, info) if return_info else x
