"""ctypes loader for oracle/_ref/librefsrc.so — the reference's OWN src/mcml_optim.cpp, src/mcml_full.cpp and src/mcml_la.cpp compiled
unmodified against oracle/shim (TEST INFRASTRUCTURE; see oracle/refsrc_driver.cpp for what is the reference's code and what is a stand-in).

The functions below have the names, argument order and return shapes of the reference's R-level exports (R/RcppExports.R) — the same as
glmmrmcml_b200's Python mirror — plus a `seed` where the reference draws random numbers (redirected to the shared Philox stream).
Built by `make -C oracle ref` where /root/reference exists; the built file travels to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "librefsrc.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
SO_OMP = os.path.join(_HERE, "_ref", "librefsrc_omp.so")   # same sources with the reference's OpenMP pragmas on: timing only (mcnr races)
_LIB = None


def available() -> bool:
    return os.path.exists(SO)


def timing_available() -> bool:
    return os.path.exists(SO_OMP)


def use_timing_build():
    """Route every call of this module to librefsrc_omp.so (bench timings; results of mcml_optim(mcnr=True) are not reproducible there)."""
    global _LIB, SO
    SO = SO_OMP
    _LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.drv_last_error.restype = C.c_char_p
        L.refsrc_set_stream.argtypes = [C.c_ulonglong, C.c_uint, C.c_long, C.c_int]
        L.refsrc_set_stream.restype = None
        _LIB = L
    return _LIB


class RefError(RuntimeError):
    """An exception thrown inside the reference's code (what R would report as an error)."""


def _check(rc):
    if rc != 0:
        raise RefError(lib().drv_last_error().decode())


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
    return (cov, data, eff), [cov.ctypes.data_as(_ip), C.c_int(cov.shape[0]), _d(data), C.c_int(data.size), _d(eff), C.c_int(eff.size)]


def _R(cov):
    npar = [0, 1, 1, 1, 2, 2, 1, 2, 2, 2, 2, 2, 2, 2, 1]           # R/R6ModelExtMCML.R:430 fnpar
    cov = np.asarray(cov, dtype=np.int32).reshape(-1, 5)
    return int(max(cov[r, 4] + npar[cov[r, 2]] for r in range(cov.shape[0])))


def mvn_ll(cov, data, eff_range, gamma, u):
    """src/mcml_optim.cpp:406-414"""
    keep, a = _cov(cov, data, eff_range)
    gamma = _v(gamma); u = _f(u if np.ndim(u) == 2 else np.reshape(u, (-1, 1)))
    out = C.c_double()
    _check(lib().drv_mvn_ll(*a, _d(gamma), C.c_int(gamma.size), _d(u), C.c_int(u.shape[0]), C.c_int(u.shape[1]), C.byref(out)))
    return out.value


def mcmc_sample(Z, L, X, y, beta, family, link, warmup, nsamp, lam, var_par=1.0, trace=0, refresh=500, maxsteps=100, target_accept=0.9, seed=1, chain=0):
    """src/mcml_full.cpp:314-338 — Q x (nsamp + 1)."""
    Z = _f(Z); L = _f(L); X = _f(X); y = _v(y); beta = _v(beta)
    n, P = X.shape; Q = Z.shape[1]
    out = np.zeros((Q, nsamp + 1), order="F")
    lib().refsrc_set_stream(seed, chain, warmup + nsamp, 0)
    _check(lib().drv_mcmc_sample(_d(Z), _d(L), _d(X), _d(y), _d(beta), C.c_int(n), C.c_int(P), C.c_int(Q), family.encode(), link.encode(), C.c_int(warmup),
                                 C.c_int(nsamp), C.c_double(lam), C.c_double(var_par), C.c_int(trace), C.c_int(refresh), C.c_int(maxsteps), C.c_double(target_accept), _d(out)))
    return out


def _fixed_u(cov, data, eff_range, Z, X, y, u):
    keep, a = _cov(cov, data, eff_range)
    Z = _f(Z); X = _f(X); y = _v(y); u = _f(u)
    n, P = X.shape; Q, m = u.shape
    return (keep, Z, X, y, u), a + [_d(Z), _d(X), _d(y), _d(u), C.c_int(n), C.c_int(P), C.c_int(Q), C.c_int(m)], (n, P, Q, m)


def mcml_optim(cov, data, eff_range, Z, X, y, u, family, link, start, trace=0, mcnr=False, _simlik=False):
    """src/mcml_optim.cpp:35-68 — dict(beta, theta, sigma)."""
    keep, a, (n, P, Q, m) = _fixed_u(cov, data, eff_range, Z, X, y, u)
    start = _v(start); R = _R(cov)
    beta = np.zeros(P); theta = np.zeros(R); sigma = C.c_double()
    _check(lib().drv_mcml_optim(*a, family.encode(), link.encode(), _d(start), C.c_int(start.size), C.c_int(trace), C.c_int(bool(mcnr)), C.c_int(bool(_simlik)),
                                _d(beta), _d(theta), C.byref(sigma)))
    return dict(beta=beta, theta=theta, sigma=sigma.value)


def mcml_simlik(cov, data, eff_range, Z, X, y, u, family, link, start, trace=0):
    """src/mcml_optim.cpp:90-117"""
    return mcml_optim(cov, data, eff_range, Z, X, y, u, family, link, start, trace, False, _simlik=True)


def mcml_hess(cov, data, eff_range, Z, X, y, u, family, link, start, tol=1e-5, trace=0):
    """src/mcml_optim.cpp:263-285 — (P + R) x (P + R)."""
    keep, a, (n, P, Q, m) = _fixed_u(cov, data, eff_range, Z, X, y, u)
    start = _v(start); k = P + _R(cov)
    H = np.zeros((k, k), order="F")
    _check(lib().drv_mcml_hess(*a, family.encode(), link.encode(), _d(start), C.c_int(start.size), C.c_double(tol), C.c_int(trace), _d(H), C.c_int(k)))
    return H


def aic_mcml(cov, data, eff_range, Z, X, y, u, family, link, beta_par, cov_par):
    """src/mcml_optim.cpp:356-392"""
    keep, a, dims = _fixed_u(cov, data, eff_range, Z, X, y, u)
    bp = _v(beta_par); cp = _v(cov_par)
    out = C.c_double()
    _check(lib().drv_aic_mcml(*a, family.encode(), link.encode(), _d(bp), C.c_int(bp.size), _d(cp), C.c_int(cp.size), C.byref(out)))
    return out.value


def mcml_full(cov, data, eff_range, Z, X, y, family, link, start, mcnr=False, m=500, maxiter=30, warmup=500, tol=1e-3, verbose=False, lam=0.05,
              trace=0, refresh=500, maxsteps=100, target_accept=0.9, seed=1):
    """src/mcml_full.cpp:41-148 — dict(beta, theta, sigma, converged, u); iteration `it` samples on the stream seed + it * 0x9E3779B97F4A7C15."""
    keep, a = _cov(cov, data, eff_range)
    Z = _f(Z); X = _f(X); y = _v(y); start = _v(start)
    n, P = X.shape; Q = Z.shape[1]; R = _R(cov)
    beta = np.zeros(P); theta = np.zeros(R); sigma = C.c_double(); conv = C.c_int(); u = np.zeros((Q, m + 1), order="F")
    lib().refsrc_set_stream(seed, 0, warmup + m, 1)
    _check(lib().drv_mcml_full(*a, _d(Z), _d(X), _d(y), C.c_int(n), C.c_int(P), C.c_int(Q), family.encode(), link.encode(), _d(start), C.c_int(start.size),
                               C.c_int(bool(mcnr)), C.c_int(m), C.c_int(maxiter), C.c_int(warmup), C.c_double(tol), C.c_double(lam), C.c_int(maxsteps),
                               C.c_double(target_accept), _d(beta), _d(theta), C.byref(sigma), C.byref(conv), _d(u)))
    return dict(beta=beta, theta=theta, sigma=sigma.value, converged=bool(conv.value), u=u)


def _la(nr, cov, data, eff_range, Z, X, y, family, link, start, usehess, tol, maxiter):
    keep, a = _cov(cov, data, eff_range)
    Z = _f(Z); X = _f(X); y = _v(y); start = _v(start)
    n, P = X.shape; Q = Z.shape[1]; R = _R(cov)
    beta = np.zeros(P); theta = np.zeros(R); sigma = C.c_double(); se = np.zeros(start.size); u = np.zeros(Q)
    _check(lib().drv_mcml_la(*a, _d(Z), _d(X), _d(y), C.c_int(n), C.c_int(P), C.c_int(Q), family.encode(), link.encode(), _d(start), C.c_int(start.size),
                             C.c_int(nr), C.c_int(bool(usehess)), C.c_double(tol), C.c_int(maxiter), _d(beta), _d(theta), C.byref(sigma), _d(se), _d(u)))
    return dict(beta=beta, theta=theta, sigma=sigma.value, se=se, u=u.reshape(Q, 1))


def mcml_la(cov, data, eff_range, Z, X, y, family, link, start, usehess=False, tol=1e-3, verbose=False, trace=0, maxiter=10):
    """src/mcml_la.cpp:28-155"""
    return _la(0, cov, data, eff_range, Z, X, y, family, link, start, usehess, tol, maxiter)


def mcml_la_nr(cov, data, eff_range, Z, X, y, family, link, start, usehess=False, tol=1e-3, verbose=False, trace=0, maxiter=10):
    """src/mcml_la.cpp:178-290"""
    return _la(1, cov, data, eff_range, Z, X, y, family, link, start, usehess, tol, maxiter)
