"""CPU oracle for the glmmrMCML hot path — TEST INFRASTRUCTURE ONLY.

ctypes loader for ``oracle/liboracle.so`` (built from ``oracle/oracle.cpp`` by ``oracle/Makefile``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; nothing under ``glmmrmcml_b200/`` does.

Parity status: *unpinned by the reference's own tests* (it has none); see the header of oracle.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """Compile liboracle.so (g++ -O2 -fopenmp) if missing or stale."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_family_ll.restype = C.c_double
        _LIB.orc_family_ll.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        _LIB.orc_log_factorial_approx.restype = C.c_double
        _LIB.orc_log_factorial_approx.argtypes = [C.c_double]
        _LIB.orc_flink.restype = C.c_int
        _LIB.orc_flink.argtypes = [C.c_char_p, C.c_char_p]
        _LIB.orc_loglik_zd.restype = C.c_double
        _LIB.orc_loglik_faithful.restype = C.c_double
        _LIB.orc_mvn_loglik.restype = C.c_double
        _LIB.orc_mcnr_sums_zd.restype = C.c_double
        _LIB.orc_logdet.restype = C.c_double
        _LIB.orc_log_prob.restype = C.c_double
        _LIB.orc_digamma.restype = C.c_double
        _LIB.orc_digamma.argtypes = [C.c_double]
        _LIB.orc_rng_uniform.restype = C.c_double
        _LIB.orc_rng_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    return _LIB


def _d(a):
    a = np.asarray(a, dtype=np.float64)
    return a.ctypes.data_as(_dp)


def _f(a):
    """column-major contiguous float64"""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _cov(cov):
    cov = np.asfortranarray(np.asarray(cov, dtype=np.int32).reshape(-1, 5))
    return cov, cov.ctypes.data_as(_ip), cov.shape[0]


def flink(family: str, link: str) -> int:
    return lib().orc_flink(family.encode(), link.encode())


def family_ll(y, mu, var_par, fl):
    return lib().orc_family_ll(float(y), float(mu), float(var_par), int(fl))


def digamma(x):
    return lib().orc_digamma(float(x))


def log_factorial_approx(n):
    return lib().orc_log_factorial_approx(float(n))


def gemm(A, B):
    A = _f(A); B = _f(B)
    M, K = A.shape; K2, N = B.shape
    assert K == K2
    Cm = np.zeros((M, N), order="F")
    lib().orc_gemm(M, N, K, _d(A), _d(B), _d(Cm))
    return Cm


def loglik_zd(zd, xb, y, var_par, fl, per_sample=False):
    """mcmlmodel.h:284-304 with zd = Z*U hoisted."""
    zd = _f(zd); n, m = zd.shape
    xb = np.ascontiguousarray(xb, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    ps = np.zeros(m)
    r = lib().orc_loglik_zd(n, m, _d(zd), _d(xb), _d(y), C.c_double(var_par), int(fl), _d(ps))
    return (r, ps) if per_sample else r


def loglik_faithful(X, Z, U, y, beta, var_par, fl, niter=None):
    """mcmlmodel.h:284-304 exactly: recomputes Z*U each call."""
    X = _f(X); Z = _f(Z); U = _f(U)
    n, P = X.shape; Q = Z.shape[1]
    niter = U.shape[1] if niter is None else niter
    y = np.ascontiguousarray(y, dtype=np.float64); beta = np.ascontiguousarray(beta, dtype=np.float64)
    return lib().orc_loglik_faithful(n, P, Q, niter, _d(X), _d(Z), _d(U), _d(y), _d(beta), C.c_double(var_par), int(fl))


def mcnr(X, Z, U, y, beta, var_par, fl, niter=None, faithful=False):
    """mcmloptim.h:198-236 (serial semantics). Returns dict(xtwx, score, beta_incr, sigma, rc)."""
    X = _f(X); Z = _f(Z); U = _f(U)
    n, P = X.shape; Q = Z.shape[1]
    niter = U.shape[1] if niter is None else niter
    y = np.ascontiguousarray(y, dtype=np.float64); beta = np.ascontiguousarray(beta, dtype=np.float64)
    xtwx = np.zeros((P, P), order="F"); score = np.zeros(P); incr = np.zeros(P); sigma = C.c_double(0)
    rc = lib().orc_mcnr(n, P, Q, niter, _d(X), _d(Z), _d(U), _d(y), _d(beta), C.c_double(var_par), int(fl),
                        int(bool(faithful)), _d(xtwx), _d(score), _d(incr), C.byref(sigma))
    return dict(xtwx=xtwx, score=score, beta_incr=incr, sigma=sigma.value, rc=rc)


def mcnr_sums_zd(zd, xb, y, var_par, fl):
    """Raw MCNR sums over the columns of zd (mcmloptim.h:210-223): (wsum[n], wusum[n], sum_j sd(resid_j)); see orc_mcnr_sums_zd."""
    zd = _f(zd); n, m = zd.shape
    xb = np.ascontiguousarray(xb, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.zeros(n); wu = np.zeros(n)
    sg = lib().orc_mcnr_sums_zd(n, m, _d(zd), _d(xb), _d(y), C.c_double(var_par), int(fl), _d(w), _d(wu))
    return w, wu, sg


def cov_dims(cov, data):
    cov, cp, rows = _cov(cov)
    data = np.ascontiguousarray(data, dtype=np.float64)
    B = C.c_int(); Q = C.c_int(); R = C.c_int()
    rc = lib().orc_cov_dims(cp, rows, _d(data), data.size, C.byref(B), C.byref(Q), C.byref(R))
    if rc:
        raise ValueError("bad covariance specification")
    return B.value, Q.value, R.value


def genD(cov, data, eff_range, theta, chol=True):
    """glmmrBase DMatrix::genD(0, chol, false): block-diagonal D or its lower Cholesky factor (Q x Q)."""
    cov, cp, rows = _cov(cov)
    data = np.ascontiguousarray(data, dtype=np.float64)
    eff = np.ascontiguousarray(eff_range if eff_range is not None and len(eff_range) else np.zeros(rows), dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    _, Q, _ = cov_dims(cov, data)
    out = np.zeros((Q, Q), order="F")
    rc = lib().orc_genD(cp, rows, _d(data), data.size, _d(eff), _d(theta), int(bool(chol)), _d(out))
    if rc:
        raise np.linalg.LinAlgError(f"D(theta) not positive definite at pivot {rc - 1}")
    return out


def mvn_loglik(cov, data, eff_range, theta, U, faithful=False):
    """MCMLDmatrix::loglik — mcmldmatrix.h:23-41 (averages ALL columns of U)."""
    cov, cp, rows = _cov(cov)
    data = np.ascontiguousarray(data, dtype=np.float64)
    eff = np.ascontiguousarray(eff_range if eff_range is not None and len(eff_range) else np.zeros(rows), dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    U = _f(np.asarray(U, dtype=np.float64).reshape(-1, 1) if np.ndim(U) == 1 else U)
    Q, m = U.shape
    return lib().orc_mvn_loglik(cp, rows, _d(data), data.size, _d(eff), _d(theta), _d(U), Q, m, int(bool(faithful)))


def logdet(cov, data, eff_range, theta):
    cov, cp, rows = _cov(cov)
    data = np.ascontiguousarray(data, dtype=np.float64)
    eff = np.ascontiguousarray(eff_range if eff_range is not None and len(eff_range) else np.zeros(rows), dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    return lib().orc_logdet(cp, rows, _d(data), data.size, _d(eff), _d(theta))


def log_prob(ZL, xb, y, var_par, fl, v):
    """mcmlmodel.h:138-153"""
    ZL = _f(ZL); n, Q = ZL.shape
    xb = np.ascontiguousarray(xb, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_log_prob(n, Q, _d(ZL), _d(xb), _d(y), C.c_double(var_par), int(fl), _d(v))


def log_grad(ZL, xb, y, var_par, fl, v):
    """mcmlmodel.h:156-279 (usezl=true)"""
    ZL = _f(ZL); n, Q = ZL.shape
    xb = np.ascontiguousarray(xb, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    g = np.zeros(Q)
    lib().orc_log_grad(n, Q, _d(ZL), _d(xb), _d(y), C.c_double(var_par), int(fl), _d(v), _d(g))
    return g


def rng_normal_vec(seed, it, chain, stream, Q):
    z = np.zeros(Q)
    lib().orc_rng_normal_vec(C.c_uint64(seed), C.c_uint32(it), C.c_uint32(chain), C.c_uint32(stream), int(Q), _d(z))
    return z


def rng_uniform(seed, it, chain, stream):
    return lib().orc_rng_uniform(C.c_uint64(seed), C.c_uint32(it), C.c_uint32(chain), C.c_uint32(stream))


def hmc_chain(ZL, L, xb, y, var_par, fl, warmup, nsamp, lam, max_steps, target_accept, seed, chain=0, adapt=100,
              want_u=True):
    """mhmcmc.h:121-157 for one chain driven by the shared Philox stream.
    Returns dict(v=Q x (nsamp+1), u=L v or None, accept, eps, ebar, steps, total_steps, prob)."""
    ZL = _f(ZL); n, Q = ZL.shape
    L = _f(L)
    xb = np.ascontiguousarray(xb, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    out_v = np.zeros((Q, nsamp + 1), order="F")
    out_u = np.zeros((Q, nsamp + 1), order="F") if want_u else None
    stats = np.zeros(5); prob = np.zeros(warmup + nsamp)
    lib().orc_hmc_chain(n, Q, _d(ZL), _d(L), _d(xb), _d(y), C.c_double(var_par), int(fl), int(warmup), int(nsamp),
                        C.c_double(lam), int(max_steps), C.c_double(target_accept), int(adapt), C.c_uint64(seed),
                        C.c_uint32(chain), _d(out_v), _d(out_u) if want_u else None, _d(stats), _d(prob))
    return dict(v=out_v, u=out_u, accept=stats[0], eps=stats[1], ebar=stats[2], steps=int(stats[3]),
                total_steps=stats[4], prob=prob)


def max_threads():
    return lib().orc_max_threads()


def set_threads(t):
    lib().orc_set_threads(int(t))
