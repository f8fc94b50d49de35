"""Synthetic inputs for the five BASELINE.json configurations.

Pure numpy; shared by tests/ and bench.py.  Everything is float64 column-major, and the covariance is
given in the reference's ``(cov, data, eff_range)`` encoding (src/mcml_optim.cpp:20-22, man/mvn_ll.Rd:10-15):
``cov`` has one row per (block, function) with columns
``[block id, block dimension, function id, number of variables, index of first parameter]``.
Function ids used here: 1 = gr, 3 = ar1, 13 = fexp (R/R6ModelExtMCML.R:430 lists the parameter counts).
"""
from __future__ import annotations

import numpy as np

SEED0 = 20221208

FN_GR, FN_FEXP0, FN_AR1, FN_SQEXP, FN_FEXP, FN_SQEXP0 = 1, 2, 3, 4, 13, 14


def dense_chol_D(cov, data, theta):
    """numpy restatement of D(theta) and its Cholesky for data generation only (not a checker)."""
    cov = np.asarray(cov).reshape(-1, 5)
    nb = int(cov[:, 0].max()) + 1
    blocks = []
    off = 0
    for b in range(nb):
        rows = cov[cov[:, 0] == b]
        n = int(rows[0, 1])
        ncol = int(rows[:, 3].sum())
        dat = np.asarray(data[off:off + n * ncol]).reshape(ncol, n).T
        off += n * ncol
        D = np.ones((n, n))
        c0 = 0
        for r in rows:
            fid, nv, p0 = int(r[2]), int(r[3]), int(r[4])
            x = dat[:, c0:c0 + nv]
            c0 += nv
            d = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
            if fid == FN_GR:
                D *= theta[p0] ** 2 * (d == 0)
            elif fid == FN_AR1:
                D *= theta[p0] ** d
            elif fid == FN_FEXP:
                D *= theta[p0] * np.exp(-d / theta[p0 + 1])
            elif fid == FN_FEXP0:
                D *= np.exp(-d / theta[p0])
            elif fid == FN_SQEXP:
                D *= theta[p0] * np.exp(-d * d / theta[p0 + 1] ** 2)
            elif fid == FN_SQEXP0:
                D *= np.exp(-d * d / theta[p0] ** 2)
            else:
                raise ValueError(fid)
        blocks.append(np.linalg.cholesky(D))
    Q = sum(b.shape[0] for b in blocks)
    L = np.zeros((Q, Q), order="F")
    s = 0
    for b in blocks:
        L[s:s + b.shape[0], s:s + b.shape[0]] = b
        s += b.shape[0]
    return L


def _finish(cfg, rng, m, u_scale=1.0):
    """draw the data y and an E-step sample matrix U = L N(0, I)."""
    L = dense_chol_D(cfg["cov"], cfg["data"], cfg["theta"])
    Q = L.shape[0]
    u_true = L @ rng.standard_normal(Q)
    eta = cfg["X"] @ cfg["beta"] + cfg["Z"] @ u_true
    fam = cfg["family"]
    if fam == "binomial":
        y = (rng.random(eta.size) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    elif fam == "poisson":
        y = rng.poisson(np.exp(eta)).astype(np.float64)
    else:
        y = eta + cfg["sigma"] * rng.standard_normal(eta.size)
    cfg["y"] = y
    cfg["L"] = L
    cfg["U"] = np.asfortranarray(u_scale * (L @ rng.standard_normal((Q, m))))
    cfg["m"] = m
    cfg["n"], cfg["P"] = cfg["X"].shape
    cfg["Q"] = Q
    cfg["eff_range"] = np.zeros(np.asarray(cfg["cov"]).reshape(-1, 5).shape[0])
    cfg["X"] = np.asfortranarray(cfg["X"])
    cfg["Z"] = np.asfortranarray(cfg["Z"])
    return cfg


def _rct_design(ncl=10, nt=5, nind=10):
    cl = np.repeat(np.arange(1, ncl + 1), nt * nind)
    t = np.tile(np.repeat(np.arange(1, nt + 1), nind), ncl)
    n = cl.size
    X = np.zeros((n, 1 + nt))
    X[:, 0] = (cl > ncl // 2)
    X[np.arange(n), t] = 1.0
    return cl, t, X


def config1(m=250, seed=SEED0 + 1, ncl=10, nt=5, nind=10):
    """C1: README cluster RCT, binomial-logit, ~(1|gr(cl)) + (1|gr(cl,t)) (README.md:14-43)."""
    rng = np.random.default_rng(seed)
    cl, t, X = _rct_design(ncl, nt, nind)
    n = cl.size
    Z = np.zeros((n, ncl + ncl * nt))
    Z[np.arange(n), cl - 1] = 1.0
    Z[np.arange(n), ncl + (cl - 1) * nt + (t - 1)] = 1.0
    cov, data = [], []
    for c in range(ncl):
        cov.append([c, 1, FN_GR, 1, 0]); data += [c + 1.0]
    b = ncl
    for c in range(ncl):
        for tt in range(nt):
            cov.append([b, 1, FN_GR, 2, 1]); data += [c + 1.0, tt + 1.0]; b += 1
    beta = np.concatenate([[0.5], rng.standard_normal(nt)])
    cfg = dict(name="C1", family="binomial", link="logit", X=X, Z=Z, cov=np.array(cov, dtype=np.int32),
               data=np.array(data), theta=np.array([0.25, 0.10]), beta=beta, sigma=1.0)
    return _finish(cfg, rng, m)


def config2(m=10_000, seed=SEED0 + 2, ncl=10, nt=5, nind=10):
    """C2: same RCT with ~(1|gr(cl)*ar1(t)) (README.md:70-77); MCNR + Hessian SEs, m = 10^4."""
    rng = np.random.default_rng(seed)
    cl, t, X = _rct_design(ncl, nt, nind)
    n = cl.size
    Z = np.zeros((n, ncl * nt))
    Z[np.arange(n), (cl - 1) * nt + (t - 1)] = 1.0
    cov, data = [], []
    for c in range(ncl):
        cov.append([c, nt, FN_GR, 1, 0]); cov.append([c, nt, FN_AR1, 1, 1])
        data += [c + 1.0] * nt + [tt + 1.0 for tt in range(nt)]
    beta = np.concatenate([[0.5], rng.standard_normal(nt)])
    cfg = dict(name="C2", family="binomial", link="logit", X=X, Z=Z, cov=np.array(cov, dtype=np.int32),
               data=np.array(data), theta=np.array([0.25, 0.8]), beta=beta, sigma=1.0)
    return _finish(cfg, rng, m)


def config3(nloc=250, m=250, seed=SEED0 + 3):
    """C3: Gaussian-identity GP, exponential covariance fexp(x,y) (README.md:104-131)."""
    rng = np.random.default_rng(seed)
    xy = rng.random((nloc, 2))
    X = np.ones((nloc, 1)); Z = np.eye(nloc)
    cov = np.array([[0, nloc, FN_FEXP, 2, 0]], dtype=np.int32)
    data = np.concatenate([xy[:, 0], xy[:, 1]])
    cfg = dict(name="C3", family="gaussian", link="identity", X=X, Z=Z, cov=cov, data=data,
               theta=np.array([0.25, 0.1]), beta=np.array([1.0]), sigma=1.0)
    return _finish(cfg, rng, m)


def config4(ncl=1000, nt=10, k=1, m=100_000, seed=SEED0 + 4):
    """C4: Poisson-log stepped-wedge trial, ncl clusters x nt periods x k obs; D = ncl ar1 blocks."""
    rng = np.random.default_rng(seed)
    cl = np.repeat(np.arange(ncl), nt * k)
    t = np.tile(np.repeat(np.arange(nt), k), ncl)
    n = cl.size
    step = 1 + (cl % (nt - 1))                   # staggered roll-out: cluster switches on at period `step`
    X = np.zeros((n, 1 + nt))
    X[:, 0] = (t >= step)
    X[np.arange(n), 1 + t] = 1.0
    Z = np.zeros((n, ncl * nt))
    Z[np.arange(n), cl * nt + t] = 1.0
    cov = np.zeros((2 * ncl, 5), dtype=np.int32)
    cov[0::2] = np.stack([np.arange(ncl), np.full(ncl, nt), np.full(ncl, FN_GR), np.ones(ncl), np.zeros(ncl)], 1)
    cov[1::2] = np.stack([np.arange(ncl), np.full(ncl, nt), np.full(ncl, FN_AR1), np.ones(ncl), np.ones(ncl)], 1)
    data = np.concatenate([np.concatenate([np.full(nt, c + 1.0), np.arange(1.0, nt + 1)]) for c in range(ncl)])
    beta = np.concatenate([[0.2], np.log(3.0) + 0.05 * rng.standard_normal(nt)])
    cfg = dict(name="C4", family="poisson", link="log", X=X, Z=Z, cov=cov, data=data,
               theta=np.array([0.25, 0.8]), beta=beta, sigma=1.0)
    return _finish(cfg, rng, m)


def config5(nloc=5000, nobs=10, m=10_000, seed=SEED0 + 5):
    """C5: large binomial-logit geospatial GLMM, one dense fexp block, nobs observations per location."""
    rng = np.random.default_rng(seed)
    xy = rng.random((nloc, 2))
    loc = np.repeat(np.arange(nloc), nobs)
    n = loc.size
    X = np.column_stack([np.ones(n), rng.standard_normal(n), rng.standard_normal(n)])
    Z = np.zeros((n, nloc))
    Z[np.arange(n), loc] = 1.0
    cov = np.array([[0, nloc, FN_FEXP, 2, 0]], dtype=np.int32)
    data = np.concatenate([xy[:, 0], xy[:, 1]])
    cfg = dict(name="C5", family="binomial", link="logit", X=X, Z=Z, cov=cov, data=data,
               theta=np.array([0.25, 0.1]), beta=np.array([-0.3, 0.4, -0.2]), sigma=1.0)
    return _finish(cfg, rng, m)


CONFIGS = {"C1": config1, "C2": config2, "C3": config3, "C4": config4, "C5": config5}
