"""glmmrmcml_b200 — B200-native Monte-Carlo E-step and random-effect sampler behind glmmrMCML's API.

Host-side mirror of the reference interface for this path.  The functions below carry the names, argument order and
return shapes of the reference's exported R functions (R/RcppExports.R:35-304): ``mcml_full``, ``mcmc_sample``,
``mcml_optim``, ``mcml_simlik``, ``mcml_hess``, ``aic_mcml``, ``mvn_ll``; ``ModelMCML`` mirrors the R6 class of
R/R6ModelExtMCML.R for the ``usestan = FALSE`` code path.  Everything calls the C-ABI of
``libglmmrmcml_b200.so`` (include/glmmrmcml_b200.h) with host numpy buffers — the same calls an Rcpp adapter makes
(INTEGRATION.md).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import GmbError, HmcStats, lib, check  # noqa: F401

__all__ = ["Context", "Model", "Covariance", "mvn_ll", "mcmc_sample", "mcml_optim", "mcml_simlik", "mcml_hess",
           "aic_mcml", "mcml_full", "mcml_la", "mcml_la_nr", "ModelMCML", "GmbError", "version"]


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _v(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _d(a):
    return a.ctypes.data_as(_lib.dp) if a is not None else None


def _covargs(cov, data, eff_range):
    cov = np.asfortranarray(np.asarray(cov, dtype=np.int32).reshape(-1, 5))
    data = _v(data)
    eff = _v(eff_range) if eff_range is not None and np.size(eff_range) else np.zeros(cov.shape[0])
    return (cov, data, eff), [cov.ctypes.data_as(_lib.ip), cov.shape[0], _d(data), data.size, _d(eff), eff.size]


def estep_set_multi(on: bool):
    """Batched binomial/logit evaluations share their pass over the factor matrix (default) or run one launch each (gmb_estep_set_multi)."""
    check(lib().gmb_estep_set_multi(int(bool(on))))


def estep_set_tf32(on: bool):
    """fp32 mode, dense Z: zd = Z u on the tensor cores (tcgen05, 3xTF32; default) or as the fp64 product narrowed to float (gmb_estep_set_tf32)."""
    check(lib().gmb_estep_set_tf32(int(bool(on))))


def hmc_set_components(on: bool):
    """Large sparse models: trajectory decomposed over the connected components of Z L (default) or one CTA per chain (gmb_hmc_set_components)."""
    check(lib().gmb_hmc_set_components(int(bool(on))))


def estep_set_sparse_zd(on: bool):
    """zd = Z u by gathering through the sparse form of Z (default, when Z is sparse) or always by the dense contraction (gmb_estep_set_sparse_zd)."""
    check(lib().gmb_estep_set_sparse_zd(int(bool(on))))


def hmc_set_lane(on: bool):
    """Small block-structured models: one lane per connected component of Z L (default) or one warp per chain (gmb_hmc_set_lane)."""
    check(lib().gmb_hmc_set_lane(int(bool(on))))


def hmc_set_factored(on: bool):
    """Two-GEMM sampler: apply a sparse Z and the dense factor L separately (default) or contract with the dense Z L (gmb_hmc_set_factored)."""
    check(lib().gmb_hmc_set_factored(int(bool(on))))


def hmc_set_variant(variant: int):
    """0 = automatic, 1 = two-GEMM sampler kernels, 2 = on-chip sampler kernel (see gmb_hmc_set_variant)."""
    check(lib().gmb_hmc_set_variant(int(variant)))


def hmc_set_cluster_size(cs: int):
    """On-chip sampler: 0 = automatic, 1 / 2 / 4 = CTAs (SMs) per group of 8 chains (see gmb_hmc_set_cluster_size)."""
    check(lib().gmb_hmc_set_cluster_size(int(cs)))


def cov_set_gram(on: bool):
    """mvn_ll on a model's samples: True (default) = Gram-matrix path for small blocks, False = stream the samples every time."""
    check(lib().gmb_cov_set_gram(int(bool(on))))


def estep_set_row_aggregation(on: bool):
    """E-step: True (default) = run on the distinct rows of [X | Z] when at most a quarter of the rows are distinct; False = every row."""
    check(lib().gmb_estep_set_row_aggregation(int(bool(on))))


def set_object_cache(on: bool):
    """Reference-named entry points: True (default) = keep the device objects of the last few models / covariance specifications between calls."""
    check(lib().gmb_set_object_cache(int(bool(on))))


def cov_set_block_classes(on: bool):
    """Gram-matrix mvn_ll: True (default) = identical covariance blocks are factorised once per class; False = once per block."""
    check(lib().gmb_cov_set_block_classes(int(bool(on))))


def hmc_set_row_aggregation(on: bool):
    """On-chip sampler: True (default) = aggregate observations that share their row of [X | Z]; False = one row per observation."""
    check(lib().gmb_hmc_set_row_aggregation(int(bool(on))))


def estep_set_rowstats(on: bool):
    """poisson / gaussian E-step: True (default) = O(n) evaluations from row statistics, False = stream zd every time."""
    check(lib().gmb_estep_set_rowstats(int(bool(on))))


def mcml_set_importance_form(reference_form: bool):
    """mcml_simlik's importance-weighted objective: False (default) = log space, True = the reference's -log(exp(ll + logl) / exp(denomD))
    (likelihood.h:101-105), which underflows for all but small models (gmb_mcml_set_importance_form)."""
    check(lib().gmb_mcml_set_importance_form(int(bool(reference_form))))


def version() -> str:
    return lib().gmb_version().decode()


class Context:
    """One per process/GPU: stream, scratch memory, optional NCCL communicator (gmb_ctx)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().gmb_ctx_create(int(device), C.byref(self._h)))
        self.rank, self.world = 0, 1

    def close(self):
        if self._h:
            lib().gmb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().gmb_ctx_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().gmb_ctx_launch_count(self._h))

    @property
    def stream(self) -> int:
        """Raw cudaStream_t of the context (integer handle), for CUDA-event timing by callers."""
        return int(lib().gmb_ctx_stream(self._h) or 0)

    def timer_start(self):
        check(lib().gmb_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        """Device time in ms since timer_start, measured with CUDA events on the context's stream."""
        ms = C.c_double()
        check(lib().gmb_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        check(lib().gmb_ctx_flush_l2(self._h))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().gmb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = C.create_string_buffer(uid, 128)
        check(lib().gmb_comm_init(self._h, buf, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def allreduce(self, a):
        a = _v(a).copy()
        check(lib().gmb_comm_allreduce_host(self._h, _d(a), a.size))
        return a

    def bcast(self, a):
        a = _v(a).copy()
        check(lib().gmb_comm_bcast_host(self._h, _d(a), a.size))
        return a

    def make_default(self):
        """Route the reference-named entry points (mcml_full, ...) through this context."""
        global _DEFAULT_CTX
        check(lib().gmb_set_default_ctx(self._h))
        _DEFAULT_CTX = self


_DEFAULT_CTX = None


def default_context() -> "Context":
    """The context behind the reference-named entry points (created on device 0 the first time it is needed)."""
    if _DEFAULT_CTX is None:
        Context(0).make_default()
    return _DEFAULT_CTX


class Model:
    """Device-resident replacement of glmmr::mcmlModel (inst/include/glmmrmcml/mcmlmodel.h:28-307)."""

    def __init__(self, ctx: Context, X, Z, y, family: str, link: str, precision: str = "fp64"):
        """precision: "fp64" (default) or "fp32" — storage of the streamed E-step matrices (gmb_model_create_prec)."""
        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        X = _f(X); Z = _f(Z); y = _v(y)
        n, P = X.shape
        if Z.shape[0] != n or y.size != n:
            raise ValueError("X, Z and y must have the same number of rows")
        self.n, self.P, self.Q = n, P, Z.shape[1]
        self.ctx = ctx
        self._h = C.c_void_p()
        self.precision = precision
        check(lib().gmb_model_create_prec(ctx._h, n, P, self.Q, _d(X), _d(Z), _d(y), family.encode(), link.encode(),
                                          32 if precision == "fp32" else 64, C.byref(self._h)))
        self.flink = lib().gmb_model_flink(self._h)

    def close(self):
        if self._h:
            if self.ctx._h:                       # a handle that outlives its context is dropped, not destroyed (the context owned its memory pool)
                lib().gmb_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_u(self, U, m_total=None, niter_total=None):
        """Upload this rank's sample columns (Q x m_local) and build zd = Z u once."""
        U = _f(np.asarray(U, dtype=np.float64).reshape(self.Q, -1))
        m_local = U.shape[1]
        m_total = m_local if m_total is None else int(m_total)
        niter_total = m_total if niter_total is None else int(niter_total)
        check(lib().gmb_model_set_u(self._h, _d(U), self.Q, m_local, m_total, niter_total))

    def use_device_u(self, niter_total=0):
        check(lib().gmb_model_use_device_u(self._h, int(niter_total)))

    def rebuild_zd(self):
        """zd = Z u again from the device-resident samples (gmb_model_rebuild_zd; timing of the contraction without the upload)."""
        check(lib().gmb_model_rebuild_zd(self._h))

    def estep_rows(self) -> int:
        """Rows of the zd matrix the model holds: n, or the number of distinct rows of [X | Z] when the E-step runs aggregated."""
        r = C.c_int()
        check(lib().gmb_model_estep_rows(self._h, C.byref(r)))
        return r.value

    def get_u(self, col0=0, ncols=None):
        """Columns [col0, col0 + ncols) of this rank's device-resident sample matrix (gmb_model_get_u)."""
        if ncols is None:
            raise ValueError("ncols is required")
        out = np.zeros((self.Q, int(ncols)), order="F")
        check(lib().gmb_model_get_u(self._h, int(col0), int(ncols), _d(out)))
        return out

    def log_likelihood(self, beta, var_par=1.0) -> float:
        """mcmlModel::log_likelihood after update_beta (mcmlmodel.h:100-102, 284-304)."""
        beta = _v(beta)
        if beta.size != self.P:
            raise ValueError(f"beta has {beta.size} values, X has {self.P} columns")
        out = C.c_double()
        check(lib().gmb_model_loglik(self._h, _d(beta), float(var_par), C.byref(out)))
        return out.value

    def log_likelihood_batch(self, betas, var_pars):
        betas = _f(np.asarray(betas, dtype=np.float64).reshape(self.P, -1))
        k = betas.shape[1]
        var_pars = _v(np.broadcast_to(np.asarray(var_pars, dtype=np.float64), (k,)))
        out = np.zeros(k)
        check(lib().gmb_model_loglik_batch(self._h, _d(betas), _d(var_pars), k, _d(out)))
        return out

    def mcnr(self, beta, var_par=1.0):
        """mcmloptim::mcnr sufficient sums and Newton step (mcmloptim.h:198-236)."""
        beta = _v(beta)
        P = self.P
        xtwx = np.zeros((P, P), order="F"); score = np.zeros(P); incr = np.zeros(P); sigma = C.c_double()
        check(lib().gmb_model_mcnr(self._h, _d(beta), float(var_par), _d(xtwx), _d(score), _d(incr), C.byref(sigma)))
        return dict(xtwx=xtwx, score=score, beta_incr=incr, sigma=sigma.value)

    def hmc_sample(self, L, beta, var_par=1.0, warmup=500, nsamp_per_chain=250, lam=5.0, max_steps=100, target_accept=0.95,
                   adapt=100, n_chains=1, chain_offset=0, seed=1, keep_on_device=False, want_u=True, want_v=False):
        """n_chains batched copies of mcmcRunHMC::sample (mhmcmc.h:121-157).  Returns dict(u, v, stats)."""
        Lh = _f(L) if L is not None else None
        beta = _v(beta)
        ncol = n_chains * (nsamp_per_chain + 1)
        U = np.zeros((self.Q, ncol), order="F") if want_u else None
        V = np.zeros((self.Q, ncol), order="F") if want_v else None
        st = HmcStats()
        check(lib().gmb_hmc_sample(self._h, _d(Lh), _d(beta), float(var_par), int(warmup), int(nsamp_per_chain), float(lam),
                                   int(max_steps), float(target_accept), int(adapt), int(n_chains), int(chain_offset),
                                   int(seed), int(bool(keep_on_device)), _d(U), _d(V), C.byref(st)))
        stats = {k: getattr(st, k) for k, _ in HmcStats._fields_}
        return dict(u=U, v=V, stats=stats)

    def log_prob_grad(self, L, beta, var_par, V):
        """mcmlModel::log_prob and log_grad (mcmlmodel.h:138-153, 156-279) for the columns of V (Q x C)."""
        Lh = _f(L) if L is not None else None
        beta = _v(beta)
        V = _f(np.asarray(V, dtype=np.float64).reshape(self.Q, -1))
        Cn = V.shape[1]
        lp = np.zeros(Cn); g = np.zeros((self.Q, Cn), order="F")
        check(lib().gmb_model_logprob_grad(self._h, _d(Lh), _d(beta), float(var_par), _d(V), Cn, _d(lp), _d(g)))
        return lp, g


class Covariance:
    """Device-resident replacement of glmmr::DData + glmmr::MCMLDmatrix (mcmldmatrix.h:18-79)."""

    def __init__(self, ctx: Context, cov, data, eff_range=None):
        self._keep, args = _covargs(cov, data, eff_range)
        self.ctx = ctx
        self._h = C.c_void_p()
        check(lib().gmb_cov_create(ctx._h, *args, C.byref(self._h)))
        B, Q, R = C.c_int(), C.c_int(), C.c_int()
        check(lib().gmb_cov_dims(self._h, C.byref(B), C.byref(Q), C.byref(R)))
        self.B, self.Q, self.R = B.value, Q.value, R.value
        nc = C.c_int()
        check(lib().gmb_cov_block_classes(self._h, C.byref(nc)))
        self.block_classes = nc.value        # distinct blocks (size, functions, data): one factorisation each on the Gram path

    def close(self):
        if self._h:
            if self.ctx._h:
                lib().gmb_cov_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def genD(self, theta, chol=True):
        """DMatrix::genD(0, chol, false): dense Q x Q block-diagonal D(theta) or its lower Cholesky factor."""
        theta = _v(theta)
        out = np.zeros((self.Q, self.Q), order="F")
        check(lib().gmb_cov_gen(self._h, _d(theta), int(bool(chol)), _d(out)))
        return out

    def loglik(self, theta, U, m_total=None) -> float:
        """MCMLDmatrix::loglik(u) (mcmldmatrix.h:23-41)."""
        theta = _v(theta)
        U = _f(np.asarray(U, dtype=np.float64).reshape(self.Q, -1))
        m = U.shape[1]
        out = C.c_double()
        check(lib().gmb_cov_mvn_ll(self._h, _d(theta), _d(U), self.Q, m, m if m_total is None else int(m_total), C.byref(out)))
        return out.value

    def loglik_model(self, theta, model: Model, ncols_total=0) -> float:
        theta = _v(theta)
        out = C.c_double()
        check(lib().gmb_cov_mvn_ll_model(self._h, _d(theta), model._h, int(ncols_total), C.byref(out)))
        return out.value

    def loglik_model_batch(self, thetas, model: Model, ncols_total=0) -> np.ndarray:
        """mvn_ll at the columns of thetas (R x k) on the model's samples: one launch, one synchronisation (-inf where D is not PD)."""
        thetas = _f(np.asarray(thetas, dtype=np.float64).reshape(self.R, -1))
        k = thetas.shape[1]
        out = np.zeros(k)
        check(lib().gmb_cov_mvn_ll_model_batch(self._h, _d(thetas), k, model._h, int(ncols_total), _d(out)))
        return out

    def logdet(self, theta) -> float:
        theta = _v(theta)
        out = C.c_double()
        check(lib().gmb_cov_logdet(self._h, _d(theta), C.byref(out)))
        return out.value


# ----------------------------------------------------------------------------------------------------------------------
# reference-named functions (R/RcppExports.R), same argument order
# ----------------------------------------------------------------------------------------------------------------------

def _fixed_u(cov, data, eff_range, Z, X, y, u, family, link):
    keep, cargs = _covargs(cov, data, eff_range)
    Z = _f(Z); X = _f(X); y = _v(y)
    n, P = X.shape
    Q = Z.shape[1]
    u = _f(np.asarray(u, dtype=np.float64).reshape(Q, -1))
    keep = keep + (Z, X, y, u)
    return keep, cargs + [_d(Z), _d(X), _d(y), _d(u), n, P, Q, u.shape[1], family.encode(), link.encode()], (n, P, Q, u.shape[1])


def cov_shape(cov):
    """(B, Q, R): blocks, total dimension and number of covariance parameters of a cov matrix (gmb_cov_shape)."""
    cov = np.asfortranarray(np.asarray(cov, dtype=np.int32).reshape(-1, 5))
    B, Q, R = C.c_int(), C.c_int(), C.c_int()
    check(lib().gmb_cov_shape(cov.ctypes.data_as(_lib.ip), cov.shape[0], C.byref(B), C.byref(Q), C.byref(R)))
    return B.value, Q.value, R.value


def _cov_R(cov):
    return cov_shape(cov)[2]


def mvn_ll(cov, data, eff_range, gamma, u) -> float:
    """src/mcml_optim.cpp:406-414."""
    keep, cargs = _covargs(cov, data, eff_range)
    gamma = _v(gamma)
    u = np.asarray(u, dtype=np.float64)
    Q = cov_shape(cov)[1]
    u = _f(u.reshape(Q, -1))
    out = C.c_double()
    check(lib().gmb_mvn_ll(*cargs, _d(gamma), gamma.size, _d(u), Q, u.shape[1], C.byref(out)))
    return out.value


def mcmc_sample(Z, L, X, y, beta, family, link, warmup, nsamp, lam, var_par=1.0, trace=0, refresh=500, maxsteps=100,
                target_accept=0.9, n_chains=1, seed=1):
    """src/mcml_full.cpp:314-338 — returns Q x (nsamp + 1) samples of u = L v."""
    Z = _f(Z); L = _f(L); X = _f(X); y = _v(y); beta = _v(beta)
    n, P = X.shape; Q = Z.shape[1]
    out = np.zeros((Q, nsamp + 1), order="F")
    check(lib().gmb_mcmc_sample(_d(Z), _d(L), _d(X), _d(y), _d(beta), n, P, Q, family.encode(), link.encode(), int(warmup),
                                int(nsamp), float(lam), float(var_par), int(trace), int(refresh), int(maxsteps),
                                float(target_accept), int(n_chains), int(seed), _d(out)))
    return out


def mcml_optim(cov, data, eff_range, Z, X, y, u, family, link, start, trace=0, mcnr=False):
    """src/mcml_optim.cpp:35-68 — list(beta, theta, sigma)."""
    keep, args, (n, P, Q, m) = _fixed_u(cov, data, eff_range, Z, X, y, u, family, link)
    start = _v(start); R = _cov_R(cov)
    beta = np.zeros(P); theta = np.zeros(R); sigma = C.c_double()
    check(lib().gmb_mcml_optim(*args, _d(start), start.size, int(trace), int(bool(mcnr)), _d(beta), _d(theta), C.byref(sigma)))
    return dict(beta=beta, theta=theta, sigma=sigma.value)


def mcml_simlik(cov, data, eff_range, Z, X, y, u, family, link, start, trace=0):
    """src/mcml_optim.cpp:90-117."""
    keep, args, (n, P, Q, m) = _fixed_u(cov, data, eff_range, Z, X, y, u, family, link)
    start = _v(start); R = _cov_R(cov)
    beta = np.zeros(P); theta = np.zeros(R); sigma = C.c_double()
    check(lib().gmb_mcml_simlik(*args, _d(start), start.size, int(trace), _d(beta), _d(theta), C.byref(sigma)))
    return dict(beta=beta, theta=theta, sigma=sigma.value)


def mcml_hess(cov, data, eff_range, Z, X, y, u, family, link, start, tol=1e-5, trace=0):
    """src/mcml_optim.cpp:263-285 — (P + R) x (P + R) finite-difference Hessian of the negative joint log-likelihood."""
    keep, args, (n, P, Q, m) = _fixed_u(cov, data, eff_range, Z, X, y, u, family, link)
    start = _v(start); R = _cov_R(cov)
    H = np.zeros((P + R, P + R), order="F")
    check(lib().gmb_mcml_hess(*args, _d(start), start.size, float(tol), int(trace), _d(H)))
    return H


def aic_mcml(cov, data, eff_range, Z, X, y, u, family, link, beta_par, cov_par) -> float:
    """src/mcml_optim.cpp:356-392."""
    keep, args, _ = _fixed_u(cov, data, eff_range, Z, X, y, u, family, link)
    beta_par = _v(beta_par); cov_par = _v(cov_par)
    out = C.c_double()
    check(lib().gmb_aic_mcml(*args, _d(beta_par), beta_par.size, _d(cov_par), cov_par.size, C.byref(out)))
    return out.value


def mcml_full(cov, data, eff_range, Z, X, y, family, link, start, mcnr=False, m=500, maxiter=30, warmup=500, tol=1e-3,
              verbose=True, lam=0.05, trace=0, refresh=500, maxsteps=100, target_accept=0.9, n_chains=0, seed=1):
    """src/mcml_full.cpp:41-148 — list(beta, theta, sigma, converged, u); u is Q x (m + 1)."""
    keep, cargs = _covargs(cov, data, eff_range)
    Z = _f(Z); X = _f(X); y = _v(y); start = _v(start)
    n, P = X.shape; Q = Z.shape[1]; R = _cov_R(cov)
    beta = np.zeros(P); theta = np.zeros(R); sigma = C.c_double(); conv = C.c_int(); it = C.c_int()
    u = np.zeros((Q, m + 1), order="F")
    check(lib().gmb_mcml_full(*cargs, _d(Z), _d(X), _d(y), n, P, Q, family.encode(), link.encode(), _d(start), start.size,
                              int(bool(mcnr)), int(m), int(maxiter), int(warmup), float(tol), int(bool(verbose)), float(lam),
                              int(trace), int(refresh), int(maxsteps), float(target_accept), int(n_chains), int(seed),
                              _d(beta), _d(theta), C.byref(sigma), C.byref(conv), C.byref(it), _d(u)))
    return dict(beta=beta, theta=theta, sigma=sigma.value, converged=bool(conv.value), iter=it.value, u=u)


def _la(fn, cov, data, eff_range, Z, X, y, family, link, start, usehess, tol, verbose, trace, maxiter):
    keep, cargs = _covargs(cov, data, eff_range)
    Z = _f(Z); X = _f(X); y = _v(y); start = _v(start)
    n, P = X.shape; Q = Z.shape[1]; R = _cov_R(cov)
    beta = np.zeros(P); theta = np.zeros(R); sigma = C.c_double(); se = np.zeros(start.size); u = np.zeros(Q); it = C.c_int()
    check(fn(*cargs, _d(Z), _d(X), _d(y), n, P, Q, family.encode(), link.encode(), _d(start), start.size, int(bool(usehess)),
             float(tol), int(bool(verbose)), int(trace), int(maxiter), _d(beta), _d(theta), C.byref(sigma), _d(se), _d(u), C.byref(it)))
    return dict(beta=beta, theta=theta, sigma=sigma.value, se=se, u=u.reshape(Q, 1), iter=it.value)


def mcml_la(cov, data, eff_range, Z, X, y, family, link, start, usehess=False, tol=1e-3, verbose=True, trace=0, maxiter=10):
    """src/mcml_la.cpp:28-155 — list(beta, theta, sigma, se, u): Laplace-approximation fit, derivative-free (beta, v) step."""
    return _la(lib().gmb_mcml_la, cov, data, eff_range, Z, X, y, family, link, start, usehess, tol, verbose, trace, maxiter)


def mcml_la_nr(cov, data, eff_range, Z, X, y, family, link, start, usehess=False, tol=1e-3, verbose=True, trace=0, maxiter=10):
    """src/mcml_la.cpp:178-290 — the same with the Newton-Raphson step mcnr_b for (beta, v)."""
    return _la(lib().gmb_mcml_la_nr, cov, data, eff_range, Z, X, y, family, link, start, usehess, tol, verbose, trace, maxiter)


def la_objectives(cov, data, eff_range, Z, X, y, family, link, beta, theta, v, sigma=1.0, w_use_l=False, newton=True):
    """Parity hook (gmb_la_objectives): the three Laplace objectives of likelihood.h:112-230 at one state and one mcnr_b step."""
    keep, cargs = _covargs(cov, data, eff_range)
    Z = _f(Z); X = _f(X); y = _v(y); beta = _v(beta); theta = _v(theta); v = _v(v)
    n, P = X.shape; Q = Z.shape[1]
    out3 = np.zeros(3); bn = np.zeros(P); vn = np.zeros(Q); sn = C.c_double()
    check(lib().gmb_la_objectives(*cargs, _d(Z), _d(X), _d(y), n, P, Q, family.encode(), link.encode(), _d(beta), _d(theta), theta.size,
                                  _d(v), float(sigma), int(bool(w_use_l)), _d(out3), _d(bn) if newton else None, _d(vn) if newton else None,
                                  C.byref(sn)))
    return dict(la=out3[0], la_cov=out3[1], la_btheta=out3[2], beta_nr=bn, v_nr=vn, sigma_nr=sn.value)


def _wrap_objective(fun, n):
    """fun maps an (n, k) array of points to k values; returns the ctypes callback (keep a reference while in use)."""
    def cb(Xp, n_, k, fp, _user):
        try:
            X = np.ctypeslib.as_array(Xp, shape=(k, n_)).T            # columns are points
            f = np.asarray(fun(X), dtype=np.float64).reshape(k)
            np.ctypeslib.as_array(fp, shape=(k,))[:] = f
            return 0
        except Exception:                                            # never let an exception cross the C boundary
            return _lib.GMB_EINVAL
    return _lib.OBJECTIVE(cb)


def minimize_bounded(fun, x0, lower=None, upper=None, rhobeg=0.0, xtol=1e-8, maxit=200):
    """gmb_minimize_bounded: batched projected-BFGS stand-in for rminqa's BOBYQA.  fun: (n, k) points -> k values."""
    x = _v(x0).copy(); n = x.size
    lo = _v(lower) if lower is not None else None
    up = _v(upper) if upper is not None else None
    cb = _wrap_objective(fun, n)
    fmin = C.c_double(); nfev = C.c_int()
    check(lib().gmb_minimize_bounded(cb, None, n, _d(x), _d(lo), _d(up), float(rhobeg), float(xtol), int(maxit), C.byref(fmin), C.byref(nfev)))
    return dict(x=x, fun=fmin.value, nfev=nfev.value)


def fd_hessian(fun, x, ndeps, lower=None, upper=None, usebounds=False):
    """gmb_fd_hessian: the optimhess stencil (4 n^2 points evaluated as one batch)."""
    x = _v(x); n = x.size
    nd = _v(np.broadcast_to(np.asarray(ndeps, dtype=np.float64), (n,)))
    lo = _v(lower) if lower is not None else None
    up = _v(upper) if upper is not None else None
    cb = _wrap_objective(fun, n)
    H = np.zeros((n, n), order="F"); nfev = C.c_int()
    check(lib().gmb_fd_hessian(cb, None, n, _d(x), _d(nd), _d(lo), _d(up), int(bool(usebounds)), _d(H), C.byref(nfev)))
    return H, nfev.value


class ModelMCML:
    """Mirror of the R6 class ``ModelMCML`` (R/R6ModelExtMCML.R).

    Holds what ``Model$new(covariance, mean.function, family)`` holds in R: the design matrices, the covariance in
    ``get_D_data()`` form, the family and starting parameters.  ``MCML(y, usestan=False)`` runs ``mcml_full``
    (R/R6ModelExtMCML.R:399-419); ``MCML(y, usestan=True)`` runs the R-level loop of :238-329 with the Stan call (``mod$sample``,
    :248-255) replaced by the native sampler (``mcmc_sample``) — sample, ``mcml_optim``, refresh the Cholesky factor — and the
    optional simulated-likelihood step (``mcml_simlik``, :335-373).  Both then take the Hessian standard errors
    (``mcml_hess``, :448-474) and the conditional AIC (``aic_mcml``, :542-553).
    """

    def __init__(self, cov, data, eff_range, Z, X, family, link, beta_start, theta_start, var_par=1.0):
        self.cov, self.data, self.eff_range = cov, data, eff_range
        self.Z, self.X = _f(Z), _f(X)
        self.family, self.link = family, link
        self.beta = _v(beta_start); self.theta = _v(theta_start); self.var_par = float(var_par)
        # R/R6ModelExtMCML.R:867-872
        self.mcmc_options = dict(warmup=500, samps=250, lam=5.0, refresh=500, maxsteps=100, target_accept=0.95)

    def chol_D(self, theta):
        """``self$covariance$get_chol_D(theta)``: dense lower Cholesky factor of D(theta)."""
        cv = Covariance(default_context(), self.cov, self.data, self.eff_range)
        try:
            return cv.genD(_v(theta), chol=True)
        finally:
            cv.close()

    def _mcml_stan_branch(self, y, start, tol, max_iter, method, sim_lik_step, verbose, n_chains, seed):
        """R/R6ModelExtMCML.R:238-373 with `mod$sample` (cmdstanr, inst/stan/*.stan: gamma ~ N(0, I), y | Xb + Z L gamma) replaced by the
        native sampler on the same target; `dsamps` = L gamma for the `samps` post-warm-up draws."""
        P = self.X.shape[1]; R = _cov_R(self.cov)
        gaussian = self.family == "gaussian"
        o = self.mcmc_options
        theta = _v(start).copy()                                  # (beta, cov pars, sigma)
        thetanew = np.ones_like(theta)                            # :182
        ib, ic, isg = slice(0, P), slice(P, P + R), P + R
        act = slice(0, P + R + 1) if gaussian else slice(0, P + R)   # all_pars, :165,:177
        L = self.chol_D(self.theta)                               # :199 get_chol_D() at the covariance's own parameters
        it = 0
        dsamps = None
        while np.any(np.abs(theta[act] - thetanew[act]) > tol) and it <= max_iter:   # :238
            it += 1
            thetanew = theta.copy()
            # C chains deliver `per` columns each, chain-major, column 0 of every chain being its state after warm-up (mhmcmc.h:142):
            # ask for exactly C * per columns so that none is truncated, and drop every chain's column 0
            Cn = max(1, min(n_chains, o["samps"]))
            per = -(-o["samps"] // Cn) + 1
            s = mcmc_sample(self.Z, L, self.X, y, thetanew[ib], self.family, self.link, o["warmup"], Cn * per - 1, o["lam"],
                            var_par=thetanew[isg], refresh=o["refresh"], maxsteps=o["maxsteps"], target_accept=o["target_accept"],
                            n_chains=Cn, seed=seed + it)
            Qd = s.shape[0]
            dsamps = np.asfortranarray(s.reshape(Qd, per, Cn, order="F")[:, 1:, :].reshape(Qd, -1, order="F")[:, :o["samps"]])   # iter_sampling draws, :256-258
            fit = mcml_optim(self.cov, self.data, self.eff_range, self.Z, self.X, y, dsamps, self.family, self.link, theta,
                             trace=0, mcnr=(method == "mcnr"))    # :293-306
            theta[ib] = fit["beta"]
            if gaussian:
                theta[isg] = fit["sigma"]
            theta[ic] = fit["theta"]
            L = self.chol_D(thetanew[ic])                         # :315 — the PREVIOUS iterate's covariance parameters, as the reference
            if verbose:
                print(f"Iter {it}: beta {theta[ib]} theta {theta[ic]} max diff {np.max(np.abs(theta[act] - thetanew[act])):.3g}")
        not_conv = it >= max_iter or bool(np.any(np.abs(theta[act] - thetanew[act]) > tol))   # :331
        if sim_lik_step:                                          # :335-373
            nt = mcml_simlik(self.cov, self.data, self.eff_range, self.Z, self.X, y, dsamps, self.family, self.link, theta)
            theta[ib] = nt["beta"]; theta[ic] = nt["theta"]
            if gaussian:
                theta[isg] = nt["sigma"]
        return dict(beta=theta[ib].copy(), theta=theta[ic].copy(), sigma=float(theta[isg]), converged=not not_conv, iter=it, u=dsamps)

    def MCML(self, y, start=None, se_theta=True, verbose=True, tol=1e-2, max_iter=30, method="mcnr", usestan=False,
             sim_lik_step=False, n_chains=0, seed=1):
        if method not in ("mcem", "mcnr"):
            raise ValueError("method must be 'mcem' or 'mcnr'")
        P = self.X.shape[1]
        if start is None:                                         # R/R6ModelExtMCML.R:159-179
            start = np.concatenate([self.beta, self.theta, [self.var_par if self.family == "gaussian" else 1.0]])
        o = self.mcmc_options
        if usestan:
            fit = self._mcml_stan_branch(y, start, tol, max_iter, method, sim_lik_step, verbose, n_chains, seed)
        else:
            fit = mcml_full(self.cov, self.data, self.eff_range, self.Z, self.X, y, self.family, self.link, start,
                            mcnr=(method == "mcnr"), m=o["samps"], maxiter=max_iter, warmup=o["warmup"], tol=tol,
                            verbose=verbose, lam=o["lam"], trace=0, refresh=o["refresh"], maxsteps=o["maxsteps"],
                            target_accept=o["target_accept"], n_chains=n_chains, seed=seed)
        out = dict(fit)
        pars = np.concatenate([fit["beta"], fit["theta"]])
        if se_theta:
            # `start = theta` in R carries sigma behind the covariance parameters (R/R6ModelExtMCML.R:464-474): f_hess holds it fixed
            hstart = np.concatenate([pars, [fit["sigma"]]]) if self.family == "gaussian" else pars
            H = mcml_hess(self.cov, self.data, self.eff_range, self.Z, self.X, y, fit["u"], self.family, self.link, hstart)
            out["hessian"] = H
            try:
                out["se"] = np.sqrt(np.diag(np.linalg.inv(H)))
            except np.linalg.LinAlgError:
                out["se"] = np.full(pars.size, np.nan)
        bp = np.concatenate([fit["beta"], [fit["sigma"]]]) if self.family == "gaussian" else fit["beta"]
        out["aic"] = aic_mcml(self.cov, self.data, self.eff_range, self.Z, self.X, y, fit["u"], self.family, self.link, bp, fit["theta"])
        return out
