"""Synthetic, already-solved problem generators for the BASELINE.json configs.

Problems are built *from* their solution (no HiGHS/Ipopt/SCS run, the primal solve is excluded
from timing anyway): pick (z, lam, nu) and data, then set h, b, q so that the KKT conditions hold
exactly.  Shared by tests/ and bench.py; no oracle or reference code involved.
"""
from __future__ import annotations

import numpy as np


def qp_batch(B, n=64, m=64, p=16, n_active=16, seed0=2026, shared=False):
    """Config 2: dense random QPs with strict complementarity and LICQ (n_active + p <= n).

    Per instance b (``default_rng(seed0 + b)``): Q = L L'/n + 0.1 I, G, A ~ N(0,1), z ~ N(0,1),
    ``n_active`` random inequality rows active (lam ~ U(.5,1.5), slack 0), the rest inactive
    (lam = 0, slack = -U(.5,1.5)), nu ~ N(0,1); h = Gz - slack, b = Az, q = -(Qz + G'lam + A'nu).
    Returns logical arrays (B, rows, cols) plus a dense forward direction and a reverse seed.
    ``shared=True``: Q, G, A identical across the batch (OptNet layer with shared weights).
    """
    d = {k: np.empty(s) for k, s in dict(
        Q=(B, n, n), G=(B, m, n), A=(B, p, n), h=(B, m), b=(B, p), q=(B, n), z=(B, n), lam=(B, m), nu=(B, p),
        dQ=(B, n, n), dq=(B, n), dG=(B, m, n), dh=(B, m), dA=(B, p, n), db=(B, p), seed=(B, n)).items()}
    for i in range(B):
        rng = np.random.default_rng(seed0 + (0 if shared else i))
        L = rng.standard_normal((n, n))
        Q = L @ L.T / n + 0.1 * np.eye(n)
        G = rng.standard_normal((m, n))
        A = rng.standard_normal((p, n))
        if shared:
            rng = np.random.default_rng(seed0 + 7919 * (i + 1))
        z = rng.standard_normal(n)
        act = rng.permutation(m)[:n_active]
        lam = np.zeros(m)
        lam[act] = rng.uniform(0.5, 1.5, size=act.size)
        slack = -rng.uniform(0.5, 1.5, size=m)
        slack[act] = 0.0
        nu = rng.standard_normal(p)
        d["Q"][i], d["G"][i], d["A"][i] = Q, G, A
        d["z"][i], d["lam"][i], d["nu"][i] = z, lam, nu
        d["h"][i] = G @ z - slack
        d["b"][i] = A @ z
        d["q"][i] = -(Q @ z + G.T @ lam + A.T @ nu)
        S = rng.standard_normal((n, n))
        d["dQ"][i] = (S + S.T) / 2
        d["dq"][i] = rng.standard_normal(n)
        d["dG"][i] = rng.standard_normal((m, n))
        d["dh"][i] = rng.standard_normal(m)
        d["dA"][i] = rng.standard_normal((p, n))
        d["db"][i] = rng.standard_normal(p)
        d["seed"][i] = rng.standard_normal(n)
    return d


def qp_batch_fast(B, n=64, m=64, p=16, n_active=16, seed=2026):
    """Same distribution as ``qp_batch`` drawn with ONE vectorised generator (for the full 4096
    batch in bench.py, where per-instance seeding only costs time)."""
    rng = np.random.default_rng(seed)
    L = rng.standard_normal((B, n, n))
    Q = L @ L.transpose(0, 2, 1) / n + 0.1 * np.eye(n)
    G = rng.standard_normal((B, m, n))
    A = rng.standard_normal((B, p, n))
    z = rng.standard_normal((B, n))
    order = np.argsort(rng.random((B, m)), axis=1)
    active = np.zeros((B, m), dtype=bool)
    np.put_along_axis(active, order[:, :n_active], True, axis=1)
    lam = np.where(active, rng.uniform(0.5, 1.5, size=(B, m)), 0.0)
    slack = np.where(active, 0.0, -rng.uniform(0.5, 1.5, size=(B, m)))
    nu = rng.standard_normal((B, p))
    h = np.einsum("bij,bj->bi", G, z) - slack
    b = np.einsum("bij,bj->bi", A, z)
    q = -(np.einsum("bij,bj->bi", Q, z) + np.einsum("bij,bi->bj", G, lam) + np.einsum("bij,bi->bj", A, nu))
    S = rng.standard_normal((B, n, n))
    return dict(Q=Q, G=G, A=A, h=h, b=b, q=q, z=z, lam=lam, nu=nu, dQ=(S + S.transpose(0, 2, 1)) / 2,
                dq=rng.standard_normal((B, n)), dG=rng.standard_normal((B, m, n)), dh=rng.standard_normal((B, m)),
                dA=rng.standard_normal((B, p, n)), db=rng.standard_normal((B, p)),
                seed=rng.standard_normal((B, n)))


def lp_config1(n=200, m=100, n_active=60, seed=1):
    """Config 1: random LP, Q == 0 -> the reference's LSQR-on-KKT branch.  Vertex-like solution:
    ``n_active`` active rows; K = [0 G'L; G D] is rank deficient by design (min-norm answer)."""
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((m, n))
    z = rng.standard_normal(n)
    act = rng.permutation(m)[:n_active]
    lam = np.zeros(m)
    lam[act] = rng.uniform(0.5, 1.5, size=n_active)
    slack = -rng.uniform(0.5, 1.5, size=m)
    slack[act] = 0.0
    return dict(Q=np.zeros((n, n)), q=-(G.T @ lam), G=G, h=G @ z - slack, A=np.zeros((0, n)), b=np.zeros(0),
                z=z, lam=lam, nu=np.zeros(0), seed=rng.standard_normal(n))


def conic_config4(n=5000, n_zero=500, n_nonneg=4000, n_soc=300, soc_dim=10, nnz_per_row=10, seed=4, psd_sides=()):
    """Config 4: sparse conic program solved by construction (Moreau decomposition):
    zeta ~ N(0,1)^m, s = Pi_K(zeta), y = s - zeta in K*, x ~ N(0,1)^n, b = A x + s, c = -A'y,
    so v = y - s = -zeta.  Returns A (scipy CSC, the reference's A = -coefficients), b, c, x, s, y,
    cone lists and a reverse seed."""
    import scipy.sparse as sp
    from oracle import cones as oc  # projection only used to *construct* the data set
    rng = np.random.default_rng(seed)
    cone_types = [oc.ZERO] * (1 if n_zero else 0) + [oc.NONNEG] * (1 if n_nonneg else 0) + [oc.SOC] * n_soc + \
        [oc.PSD] * len(psd_sides)
    cone_dims = ([n_zero] if n_zero else []) + ([n_nonneg] if n_nonneg else []) + [soc_dim] * n_soc + \
        [d * (d + 1) // 2 for d in psd_sides]
    m = int(sum(cone_dims))
    rows = np.repeat(np.arange(m), nnz_per_row)
    cols = rng.integers(0, n, size=m * nnz_per_row)
    vals = rng.standard_normal(m * nnz_per_row)
    A = sp.csc_matrix((vals, (rows, cols)), shape=(m, n))
    A.sum_duplicates()
    A.sort_indices()
    zeta = rng.standard_normal(m)
    # primal cone K = product of the MOI sets; s = Pi_K(zeta).  oracle.cones projects on the DUAL of
    # the listed set, which equals the set itself for nonneg/SOC/PSD; for Zeros, K = {0}.
    s = np.empty(m)
    off = oc.cone_offsets(cone_dims)
    for k, t in enumerate(cone_types):
        sl = slice(off[k], off[k + 1])
        s[sl] = 0.0 if t == oc.ZERO else oc.project(zeta[sl], t)
    y = s - zeta
    x = rng.standard_normal(n)
    return dict(A=A, b=A @ x + s, c=-(A.T @ y), x=x, s=s, y=y, cone_types=cone_types, cone_dims=cone_dims,
                seed=rng.standard_normal(n))
