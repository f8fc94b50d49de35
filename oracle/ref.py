"""ctypes loader for oracle/_ref/libref.so — the reference's own headers compiled against oracle/shim (TEST INFRASTRUCTURE).

Built by `make -C oracle ref` (only possible where /root/reference exists; the built file travels to the GPU box).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref.so")
SO_OMP = os.path.join(_HERE, "_ref", "libref_omp.so")   # same sources with the reference's OpenMP pragmas on: timing only
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_LIB = None


def available() -> bool:
    return os.path.exists(SO)


def timing_available() -> bool:
    return os.path.exists(SO_OMP)


def use_timing_build():
    """Route every call of this module to libref_omp.so (bench.py's reference arm; results of mcnr are not reproducible there)."""
    global _LIB, SO
    SO = SO_OMP
    _LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        for nm in ("ref_family_ll", "ref_log_factorial_approx", "ref_loglik", "ref_loglik_reps", "ref_log_prob", "ref_mvn_loglik", "ref_logdet"):
            getattr(L, nm).restype = C.c_double
        L.ref_family_ll.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        L.ref_log_factorial_approx.argtypes = [C.c_double]
        L.ref_root.restype = C.c_char_p
        _LIB = L
    return _LIB


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _v(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _d(a):
    return a.ctypes.data_as(_dp)


def _cov(cov, data, eff):
    cov = np.asfortranarray(np.asarray(cov, dtype=np.int32).reshape(-1, 5))
    data = _v(data)
    eff = _v(eff) if eff is not None and np.size(eff) else np.zeros(cov.shape[0])
    return (cov, data, eff), [cov.ctypes.data_as(_ip), cov.shape[0], _d(data), data.size, _d(eff), eff.size]


def family_ll(y, mu, var_par, fl):
    return lib().ref_family_ll(float(y), float(mu), float(var_par), int(fl))


def log_factorial_approx(n):
    return lib().ref_log_factorial_approx(float(n))


def detadmu(xb, link):
    xb = _v(xb); out = np.zeros(xb.size)
    lib().ref_detadmu(_d(xb), xb.size, link.encode(), _d(out))
    return out


def forward_sub(L, u):
    L = _f(L); u = _v(u); out = np.zeros(u.size)
    lib().ref_forward_sub(_d(L), _d(u), u.size, _d(out))
    return out


def loglik(X, Z, U, y, beta, var_par, family, link):
    X = _f(X); Z = _f(Z); U = _f(U); y = _v(y); beta = _v(beta)
    n, P = X.shape; Q, m = U.shape
    return lib().ref_loglik(n, P, Q, m, _d(X), _d(Z), _d(U), _d(y), _d(beta), C.c_double(var_par), family.encode(), link.encode())


def loglik_reps(X, Z, U, y, beta, var_par, family, link, reps):
    """model built once, log_likelihood() evaluated `reps` times (timing helper)."""
    X = _f(X); Z = _f(Z); U = _f(U); y = _v(y); beta = _v(beta)
    n, P = X.shape; Q, m = U.shape
    return lib().ref_loglik_reps(n, P, Q, m, _d(X), _d(Z), _d(U), _d(y), _d(beta), C.c_double(var_par), family.encode(), link.encode(), int(reps))


def log_prob(X, Z, L, y, beta, var_par, family, link, v):
    X = _f(X); Z = _f(Z); L = _f(L); y = _v(y); beta = _v(beta); v = _v(v)
    n, P = X.shape; Q = Z.shape[1]
    return lib().ref_log_prob(n, P, Q, _d(X), _d(Z), _d(L), _d(y), _d(beta), C.c_double(var_par), family.encode(), link.encode(), _d(v))


def log_grad(X, Z, L, y, beta, var_par, family, link, v):
    X = _f(X); Z = _f(Z); L = _f(L); y = _v(y); beta = _v(beta); v = _v(v)
    n, P = X.shape; Q = Z.shape[1]
    g = np.zeros(Q)
    lib().ref_log_grad(n, P, Q, _d(X), _d(Z), _d(L), _d(y), _d(beta), C.c_double(var_par), family.encode(), link.encode(), _d(v), _d(g))
    return g


def mvn_loglik(cov, data, eff, theta, U):
    keep, a = _cov(cov, data, eff)
    theta = _v(theta); U = _f(np.asarray(U, dtype=np.float64).reshape(-1, 1) if np.ndim(U) == 1 else U)
    Q, m = U.shape
    return lib().ref_mvn_loglik(*a, _d(theta), theta.size, _d(U), Q, m)


def logdet(cov, data, eff, theta):
    keep, a = _cov(cov, data, eff)
    theta = _v(theta)
    return lib().ref_logdet(*a, _d(theta), theta.size)


def mcnr(cov, data, eff, X, Z, U, y, family, link, start):
    """mcml_optim(..., mcnr=TRUE) up to mc.mcnr(): returns (beta after the Newton step, sigma)."""
    keep, a = _cov(cov, data, eff)
    X = _f(X); Z = _f(Z); U = _f(U); y = _v(y); start = _v(start)
    n, P = X.shape; Q, m = U.shape
    beta = np.zeros(P); sigma = C.c_double()
    lib().ref_mcnr(*a, n, P, Q, m, _d(X), _d(Z), _d(U), _d(y), family.encode(), link.encode(), _d(start), start.size, _d(beta), C.byref(sigma))
    return beta, sigma.value


def objectives(cov, data, eff, X, Z, U, y, family, link, par, fix_var_par=1.0):
    """(L_likelihood, D_likelihood, F_likelihood[importance=false, fix_var=true]) at par = (beta, theta)."""
    keep, a = _cov(cov, data, eff)
    X = _f(X); Z = _f(Z); U = _f(U); y = _v(y); par = _v(par)
    n, P = X.shape; Q, m = U.shape
    out = np.zeros(3)
    lib().ref_objectives(*a, n, P, Q, m, _d(X), _d(Z), _d(U), _d(y), family.encode(), link.encode(), _d(par), par.size, C.c_double(fix_var_par), _d(out))
    return out


def la_objectives(cov, data, eff, X, Z, y, family, link, beta, theta, v, sigma=1.0, w_use_l=False):
    """LA_likelihood / LA_likelihood_cov / LA_likelihood_btheta (likelihood.h:112-230) at one state + one mcnr_b step."""
    keep, a = _cov(cov, data, eff)
    X = _f(X); Z = _f(Z); y = _v(y); beta = _v(beta); theta = _v(theta); v = _v(v)
    n, P = X.shape; Q = Z.shape[1]
    out3 = np.zeros(3); bn = np.zeros(P); vn = np.zeros(Q); sn = C.c_double()
    lib().ref_la_objectives(*a, n, P, Q, _d(X), _d(Z), _d(y), family.encode(), link.encode(), _d(beta), _d(theta), theta.size, _d(v),
                            C.c_double(sigma), int(bool(w_use_l)), _d(out3), _d(bn), _d(vn), C.byref(sn))
    return dict(la=out3[0], la_cov=out3[1], la_btheta=out3[2], beta_nr=bn, v_nr=vn, sigma_nr=sn.value)


def mcmc_sample(X, Z, L, y, beta, family, link, warmup, nsamp, lam, var_par, maxsteps, target_accept, seed, chain=0):
    """mcmc_sample (src/mcml_full.cpp:314-338) driven by the shared Philox stream; returns (Q x (nsamp+1) u-samples, stats)."""
    X = _f(X); Z = _f(Z); L = _f(L); y = _v(y); beta = _v(beta)
    n, P = X.shape; Q = Z.shape[1]
    out = np.zeros((Q, nsamp + 1), order="F"); st = np.zeros(4)
    lib().ref_mcmc_sample(n, P, Q, _d(X), _d(Z), _d(L), _d(y), _d(beta), family.encode(), link.encode(), int(warmup), int(nsamp),
                          C.c_double(lam), C.c_double(var_par), int(maxsteps), C.c_double(target_accept), C.c_uint64(seed),
                          C.c_uint32(chain), _d(out), _d(st))
    return out, dict(accept=st[0], eps=st[1], ebar=st[2], steps=int(st[3]))
