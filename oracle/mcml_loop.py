"""Oracle-side restatement of the MCML loop of src/mcml_full.cpp:62-146 — TEST INFRASTRUCTURE ONLY.

One sequential HMC chain per iteration (oracle.hmc_chain = mhmcmc.h:121-157, re-initialised and re-adapted every iteration, :127), the
beta step as MCNR (mcmloptim.h:198-236) or as a bounded minimisation of the Monte-Carlo log-likelihood (l_optim, :71-88), the theta step
as a bounded minimisation of the multivariate-normal objective (d_optim, :56-68; theta >= 1e-6), the reference's convergence test
(src/mcml_full.cpp:108-113) and its quirk that u has m + 1 columns while the E-step uses niter_ = m of them.

rminqa's BOBYQA is not available offline; scipy's bounded minimisers stand in for it (any correct bounded optimiser reaches the same
optimum to far below the MCML tolerance — the tests compare at 1e-4 against tol = 5e-3 / 1e-2).

The chain of iteration `it` is keyed by seed + it * 0x9E3779B97F4A7C15 (mod 2^64) on the shared Philox stream, which is what
gmb_mcml_full uses, so that with n_chains = 1 the device loop can be followed iterate by iterate."""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize

from . import flink as _flink, genD, gemm, hmc_chain, loglik_zd, mcnr as _mcnr, mvn_loglik

GOLDEN = 0x9E3779B97F4A7C15
MASK = (1 << 64) - 1


def _min_bounded(f, x0, lower):
    """Bounded minimiser standing in for Rbobyqa: L-BFGS-B, then a Nelder-Mead polish from its optimum (projected onto the bounds)."""
    x0 = np.asarray(x0, dtype=np.float64)
    lo = np.asarray(lower, dtype=np.float64)
    g = lambda x: f(np.maximum(x, lo)) if np.all(np.isfinite(x)) else 1e300
    r = minimize(g, x0, method="L-BFGS-B", bounds=[(l if np.isfinite(l) else None, None) for l in lo], options=dict(ftol=1e-15, gtol=1e-10, maxiter=500))
    r2 = minimize(g, r.x, method="Nelder-Mead", options=dict(xatol=1e-10, fatol=1e-15, maxiter=4000, initial_simplex=None))
    x = np.maximum(r2.x if r2.fun <= r.fun else r.x, lo)
    return x


def mcml_full(cov, data, eff_range, Z, X, y, family, link, start, mcnr=False, m=500, maxiter=30, warmup=500, tol=1e-3, lam=0.05,
              maxsteps=100, target_accept=0.9, seed=1, trace=None):
    """src/mcml_full.cpp:41-148 with the oracle's pieces.  Returns dict(beta, theta, sigma, converged, iter, u, path)."""
    X = np.asfortranarray(X, dtype=np.float64); Z = np.asfortranarray(Z, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    start = np.asarray(start, dtype=np.float64)
    P = X.shape[1]
    fl = _flink(family, link)
    R = start.size - P - 1
    beta = start[:P].copy(); theta = start[P:P + R].copy()                       # :63-64
    has_var = family in ("gaussian", "Gamma")
    var_par = float(start[-1]) if has_var else 1.0                               # :65
    sigma = var_par if family == "gaussian" else 0.0                             # mcmloptim ctor, mcmloptim.h:30
    L = genD(cov, data, eff_range, theta, chol=True)                             # :68
    it, maxdiff, converged = 1, 1.0, False
    path = []
    U = None
    while maxdiff > tol and it <= maxiter:                                       # :83
        ZL = gemm(Z, L)
        ch = hmc_chain(ZL, L, X @ beta, y, var_par, fl, warmup, m, lam, maxsteps, target_accept, (seed + it * GOLDEN) & MASK, chain=0)   # :92
        U = ch["u"]                                                              # Q x (m + 1); niter_ = m
        zd = gemm(Z, np.asfortranarray(U[:, :m]))
        if mcnr:                                                                 # :98 -> mcmloptim.h:198-236
            r = _mcnr(X, Z, U, y, beta, var_par, fl, niter=m)
            newbeta = beta + r["beta_incr"]; sigma = r["sigma"]
        else:                                                                    # :96 -> mcmloptim.h:71-88
            if family == "gaussian":
                f = lambda p: -loglik_zd(zd, X @ p[:P], y, p[P], fl) if p[P] > 0 else 1e300
                x = _min_bounded(f, np.concatenate([beta, [sigma]]), np.concatenate([np.full(P, -np.inf), [0.0]]))
                newbeta, sigma = x[:P], float(x[P])
            else:
                f = lambda p: -loglik_zd(zd, X @ p, y, 0.0, fl)
                newbeta = _min_bounded(f, beta, np.full(P, -np.inf))

        def dobj(th):                                                            # D_likelihood, likelihood.h:40-45 over ALL m + 1 columns
            v = mvn_loglik(cov, data, eff_range, th, U)
            return -v if np.isfinite(v) else 1e300
        newtheta = _min_bounded(dobj, theta, np.full(R, 1e-6))                   # :101 -> mcmloptim.h:56-68
        new_var_par = sigma if has_var else var_par                              # :105 (new_var_par starts at 1, :79)
        if not has_var:
            new_var_par = 1.0
        maxdiff = max(np.max(np.abs(beta - newbeta)), np.max(np.abs(theta - newtheta)), abs(var_par - new_var_par))   # :108-111
        if maxdiff < tol:
            converged = True                                                     # :113
        beta, theta, var_par = newbeta.copy(), newtheta.copy(), new_var_par      # :116-118
        path.append(dict(iter=it, beta=beta.copy(), theta=theta.copy(), sigma=var_par, maxdiff=float(maxdiff), accept=ch["accept"], eps=ch["eps"]))
        if trace:
            trace(path[-1])
        if not converged:
            L = genD(cov, data, eff_range, theta, chol=True)                     # :119-126
        it += 1
    return dict(beta=beta, theta=theta, sigma=var_par, converged=converged, iter=it - 1, u=U, path=path)
