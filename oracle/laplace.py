"""oracle/laplace.py — numpy restatement of the Laplace-approximation path of the reference (TEST INFRASTRUCTURE).

Follows inst/include/glmmrmcml/likelihood.h:112-230 (LA_likelihood, LA_likelihood_cov, LA_likelihood_btheta),
mcmloptim.h:116-195, 238-293 (la_optim, la_optim_cov, la_optim_bcov, hess_la, mcnr_b), mcmlmodel.h:120-134 (update_W) and the
drivers src/mcml_la.cpp:28-155, 178-290.  Small dense linear algebra only, so numpy is the checker; the per-observation
family terms and D(theta) come from the C++ oracle (oracle.cpp).  Only tests/ may import this module.

Quirks kept on purpose (each is what the reference computes):
  * the state optimised by la_optim / mcnr_b is the WHITENED vector v (u = L v is formed only for the return value, src/mcml_la.cpp:148);
    the prior term is v'v/2 (likelihood.h:126,161,206)
  * update_W() without arguments forms the weights at eta = xb + Z v — Z, not Z L (mcmlmodel.h:121, useL = false); mcml_la and
    LA_likelihood_btheta use that form, mcml_la_nr calls update_W(0, true) (eta = xb + Z L v)
  * mcnr_b evaluates mu at xb + Z L v but takes log_grad(v, usezl = false): mu = xb + Z v, grad = -D v + (Z L)' r(mu), with D = L L' of
    the INITIAL theta (update_D() is commented out at src/mcml_la.cpp:248)
  * sigma: var_par_ starts at 1 (src/mcml_la.cpp:46,194) whatever `start` holds
"""
from __future__ import annotations

import numpy as np

import oracle as _o


def inv_link(eta, link):
    """glmmrBase maths::mod_inv_func (reconstructed)."""
    if link == "logit":
        return np.exp(eta) / (1 + np.exp(eta))
    if link == "log":
        return np.exp(eta)
    return np.asarray(eta, dtype=np.float64).copy()


def dhdmu(eta, family, link):
    """glmmrBase maths::dhdmu (reconstructed): reciprocal IRLS weight without the dispersion."""
    if family == "poisson" and link == "log":
        return np.exp(-eta)
    if family == "binomial" and link == "logit":
        p = inv_link(eta, "logit")
        return 1 / (p * (1 - p))
    return np.ones_like(eta)


def detadmu(eta, link):
    """moremaths.h:118-161."""
    if link == "log":
        return np.exp(-1.0 * eta)
    if link == "logit":
        p = inv_link(eta, "logit")
        return 1 / (p * (1.0 - p))
    return np.ones_like(eta)


def w_diag(xb, zu, var_par, family, link):
    """mcmlModel::update_W, mcmlmodel.h:120-134: diagonal of W_."""
    w = dhdmu(xb + zu, family, link)
    nvar = var_par * var_par if family == "gaussian" else 1.0
    return 1 / (w * nvar)


def _ll_sum(y, eta, var_par, fl):
    """sum_i maths::log_likelihood(y_i, eta_i, var_par, flink) (moremaths.h:26-102) through the C++ oracle."""
    return _o.loglik_zd(np.asfortranarray(eta.reshape(-1, 1)), np.zeros(eta.size), y, var_par, fl)


def logdet_llt(M):
    """glmmr::maths::logdet, moremaths.h:104-116."""
    return 2 * np.sum(np.log(np.diag(np.linalg.cholesky(M))))


def la_likelihood(par, X, ZL, y, var_par, fl):
    """LA_likelihood::operator(), likelihood.h:121-140: par = (beta, v)."""
    P = X.shape[1]
    beta, v = par[:P], par[P:]
    logl = v @ v
    ll = _ll_sum(y, X @ beta + ZL @ v, var_par, fl)
    return -1.0 * (ll - 0.5 * logl)


def la_likelihood_cov(par, cov, data, eff, Z, xb, y, v, W, family, fl, var_par):
    """LA_likelihood_cov::operator(), likelihood.h:153-180: par = theta (+ sigma for gaussian)."""
    has_s = family == "gaussian"
    theta = np.asarray(par[:-1] if has_s else par, dtype=np.float64)
    if has_s:
        var_par = par[-1]
    L = _o.genD(cov, data, eff, theta, chol=True)
    ZL = Z @ L
    ll = _ll_sum(y, xb + ZL @ v, var_par, fl)
    M = ZL.T @ (W[:, None] * ZL) + np.eye(L.shape[0])
    return -1 * (ll - 0.5 * (v @ v) - 0.5 * logdet_llt(M))


def la_likelihood_btheta(par, cov, data, eff, Z, X, y, v, family, link, fl, var_par):
    """LA_likelihood_btheta::operator(), likelihood.h:193-229: par = (beta, theta[, sigma]); W is refreshed with update_W() (Z v)."""
    P = X.shape[1]
    has_s = family == "gaussian"
    R = len(par) - P - (1 if has_s else 0)
    beta = np.asarray(par[:P]); theta = np.asarray(par[P:P + R], dtype=np.float64)
    if has_s:
        var_par = par[-1]
    xb = X @ beta
    W = w_diag(xb, Z @ v, var_par, family, link)                     # update_W(): useL = false
    return la_likelihood_cov(np.concatenate([theta, [var_par]]) if has_s else theta, cov, data, eff, Z, xb, y, v, W, family, fl, var_par)


def mcnr_b(X, Z, L, D0, y, beta, v, W, var_par, family, link, fl):
    """mcmloptim::mcnr_b, mcmloptim.h:238-293.  Returns (beta + bincr, v + vincr, sigma)."""
    ZL = Z @ L
    xb = X @ beta
    zd = ZL @ v
    dmu = detadmu(xb + zd, link)
    M = ZL.T @ (W[:, None] * ZL) + np.eye(L.shape[0])
    Minv = np.linalg.solve(M, np.eye(L.shape[0]))                    # LZWZL.llt().solve(I)
    resid = y - inv_link(xb + zd, link)
    sigmas = np.sqrt(np.sum((resid - resid.mean()) ** 2) / (resid.size - 1))
    Wu = W * dmu * resid
    XtWX = np.linalg.inv(X.T @ (W[:, None] * X))
    bincr = XtWX @ (X.T @ Wu)
    # log_grad(v, usezl = false), mcmlmodel.h:156-279: mu = xb + Z v, grad = -D v + (Z L)' r(mu)
    mu = xb + Z @ v
    if fl == 1:
        r = y - np.exp(mu)
    elif fl == 3:
        r = 1 / (np.exp(mu) + 1) + y - 1
    else:
        r = (y - mu) / (var_par * var_par)
    vgrad = -1.0 * (D0 @ v) + ZL.T @ r
    vincr = Minv @ vgrad
    return beta + bincr, v + vincr, sigmas
