#!/usr/bin/env python
"""BASELINE.json's large configurations at their STATED size, each with an in-run parity check against the CPU oracle.

  C3  Gaussian-identity GP, exponential covariance, n = Q = 10^4 locations (one dense 10^4 x 10^4 block), m = 250
  C4  Poisson-log stepped wedge, 1000 clusters x 10 periods (n = Q = 10^4, 1000 ar1 blocks of 10), m = 10^5 samples in total
  C5  binomial-logit geospatial, n = 5 10^4, Q = 5 10^3 (one dense block), m = 10^4, 1024 chains in total

Multi-GPU: the m sample columns and the chains are split over the ranks (STRONG scaling: total work fixed); C3 (m = 250) does not
shard and runs at N = 1 only.  Per configuration and kernel family the block reports the throughput BASELINE.json's metric names
(E-step log-likelihood evaluations/s, MCNR steps/s, mvn_ll evaluations/s, sampler u-samples/s and leapfrog steps/s), the roofline
fraction against the bound SURVEY.md §8(d) names (algorithmic bytes / flops of §8(d) per launch / CUDA-event time of the launch, per GPU),
and a `parity` object: the GPU value against the oracle on THE SAME inputs —

  loglik   oracle.loglik_zd (mcmlmodel.h:284-304), walked in column chunks              relative 1e-10
  mcnr     oracle.mcnr_sums_zd (mcmloptim.h:198-236): X'WX, score, sigma                  relative 1e-10 (max-norm)
  mvn_ll   oracle.mvn_loglik (mcmldmatrix.h:23-78) for small blocks; for one large dense block the same formula on LAPACK
           (numpy Cholesky + scipy triangular solve — the oracle's plain loops need minutes at Q = 10^4)   relative 1e-10
  chol     device factor against LAPACK dpotrf, max-norm, in units of kappa_1(D) eps (dpocon)             <= 1e-12 kappa / eps-scaled
  sampler  log_prob / log_grad (mcmlmodel.h:138-279) at random states (1e-10) and chain 0 state by state against the oracle's
           mcmcRunHMC chain under the same Philox stream (1e-7), on the kernel family the dispatcher picks at this size

Used by bench.py (the `configs` block of its JSON line), by tests/test_gpu_fullsize.py and from the command line:
  python tools/bench_configs.py [C3 C4 C5] [--small]
Only this file, bench.py and tests/ touch the oracle; the product package never does.
"""
from __future__ import annotations

import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FP64_PEAK_TFLOPS = 37.1      # measured DMMA / DFMA pipe peak of this pool's B200 (profiles/r01_microbench_fp64.txt)
SIZES = {
    "full": {"C3": dict(nloc=10_000, m=250, chains=256, hmc=dict(warmup=4, nsamp=1)),
             "C4": dict(ncl=1000, nt=10, m=100_000, chains=1024, hmc=dict(warmup=100, nsamp=9)),
             "C5": dict(nloc=5000, nobs=10, m=10_000, chains=1024, hmc=dict(warmup=8, nsamp=3))},
    "small": {"C3": dict(nloc=1500, m=96, chains=64, hmc=dict(warmup=4, nsamp=1)),
              "C4": dict(ncl=120, nt=10, m=4096, chains=128, hmc=dict(warmup=20, nsamp=3)),
              "C5": dict(nloc=700, nobs=6, m=1024, chains=128, hmc=dict(warmup=6, nsamp=2))},
}
HMC_DEFAULTS = dict(lam=5.0, max_steps=100, target_accept=0.95)       # R/R6ModelExtMCML.R:867-872


# ----------------------------------------------------------------------------------------------------------------------
# host-side data
# ----------------------------------------------------------------------------------------------------------------------
def fast_normal(ncols, nrows, seed, threads=None):
    """(ncols, nrows) C-ordered standard normals — column j of the (nrows x ncols) Fortran view is contiguous.  numpy's generators
    release the GIL while filling, so independent streams fill disjoint row blocks from a thread pool."""
    threads = threads or min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 8)
    out = np.empty((ncols, nrows))
    bounds = np.linspace(0, ncols, threads + 1).astype(int)

    def fill(t):
        a, b = bounds[t], bounds[t + 1]
        if b > a:
            np.random.default_rng([seed, t]).standard_normal(out=out[a:b])
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(fill, range(threads)))
    return out


def indicator_index(Z):
    """Row -> column of the single 1 when Z is an indicator matrix with exactly one non-zero per row, else None."""
    nz = np.count_nonzero(Z, axis=1)
    if np.all(nz == 1):
        idx = np.argmax(Z != 0, axis=1)
        if np.all(Z[np.arange(Z.shape[0]), idx] == 1.0):
            return idx
    return None


def build(name, size):
    """The model of configuration `name` (every rank builds the same one) and what is needed to draw sample columns."""
    from glmmrmcml_b200 import synth
    p = SIZES[size][name]
    if name == "C3":
        cfg = synth.config3(nloc=p["nloc"], m=4)
    elif name == "C4":
        cfg = synth.config4(ncl=p["ncl"], nt=p["nt"], k=1, m=4)
    else:
        cfg = synth.config5(nloc=p["nloc"], nobs=p["nobs"], m=4)
    cfg["zidx"] = indicator_index(cfg["Z"])
    cfg["m_total"] = p["m"]
    cfg["chains_total"] = p["chains"]
    cfg["hmc"] = dict(HMC_DEFAULTS, **p["hmc"])
    return cfg


def draw_u(cfg, name, ncols, seed):
    """u = L z for `ncols` columns, Q x ncols column-major."""
    Q = cfg["Q"]
    z = fast_normal(ncols, Q, seed)                       # (ncols, Q): row j = z_j'
    if name == "C4":                                      # identical 10 x 10 blocks: one small GEMM on the (ncols * ncl, nt) view
        nt = int(np.asarray(cfg["cov"]).reshape(-1, 5)[0, 1])
        Lb = cfg["L"][:nt, :nt]
        u = (z.reshape(-1, nt) @ Lb.T).reshape(ncols, Q)
    else:
        u = z @ cfg["L"].T
    return u.T                                            # Fortran-ordered (Q, ncols) view


def z_apply(cfg, Uc):
    """zd = Z u for a block of columns (gather when Z is an indicator)."""
    return np.asfortranarray(Uc[cfg["zidx"], :]) if cfg["zidx"] is not None else np.asfortranarray(cfg["Z"] @ Uc)


def col_chunks(m, n_rows, budget_doubles=2.5e8):
    step = max(1, int(budget_doubles // max(1, n_rows)))
    return [(a, min(m, a + step)) for a in range(0, m, step)]


# ----------------------------------------------------------------------------------------------------------------------
class Env:
    """What run_config needs from its caller: the package, a context (possibly joined to NCCL), rank / world and reductions."""

    def __init__(self, g, ctx, rank=0, world=1, dist=None, hbm_gbs=6650.0):
        self.g, self.ctx, self.rank, self.world, self.dist, self.hbm = g, ctx, rank, world, dist, hbm_gbs

    def _red(self, x, op):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor(np.atleast_1d(np.asarray(x, dtype=np.float64)), device="cuda")
        self.dist.all_reduce(t, op=op)
        r = t.cpu().numpy()
        return float(r[0]) if np.ndim(x) == 0 else r

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.dist else None)

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.dist else None)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b)) / d) if d > 0 else float(np.max(np.abs(a - b)))


def timed(ctx, fn, reps=3, flush=False):
    ts = []
    for _ in range(reps):
        if flush:
            ctx.flush_l2()
        ctx.sync()
        ctx.timer_start(); fn(); ts.append(ctx.timer_stop())
    return float(np.median(ts))


def run_config(name, env, size="full", oracle=None, check=True):
    """One configuration: throughput, roofline fractions and parity.  Collective on every rank of env."""
    g, ctx, rank, world = env.g, env.ctx, env.rank, env.world
    if oracle is None:
        import oracle as oracle_mod
        oracle = oracle_mod
        oracle.build()
    t_host0 = time.perf_counter()
    cfg = build(name, size)
    n, P, Q = cfg["n"], cfg["P"], cfg["Q"]
    fam, link, beta, theta, sig = cfg["family"], cfg["link"], cfg["beta"], cfg["theta"], cfg.get("sigma", 1.0)
    fl = oracle.flink(fam, link)
    m_total = cfg["m_total"]
    lo, hi = rank * m_total // world, (rank + 1) * m_total // world
    m_local = hi - lo
    U = draw_u(cfg, name, m_local, 7_000_000 + 1000 * rank + int(name[1]))
    out = {"config": name, "size": size, "n": n, "P": P, "Q": Q, "m_total": m_total, "m_per_gpu": m_local, "family": fam,
           "blocks": int(np.asarray(cfg["cov"]).reshape(-1, 5)[:, 0].max()) + 1, "ranks": world,
           "host_setup_s": round(time.perf_counter() - t_host0, 2)}
    parity = {}
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], fam, link)
    cv = g.Covariance(ctx, cfg["cov"], cfg["data"], cfg["eff_range"])
    large_block = cv.B == 1 and cv.Q > 64

    # ---- K1: upload of this rank's columns + zd = Z u -----------------------------------------------------------------
    t0 = time.perf_counter(); mdl.set_u(U, m_total=m_total); out["set_u_s"] = env.max(time.perf_counter() - t0)
    out["zd_bytes_per_gpu"] = 8.0 * n * m_local

    # ---- K2: E-step objective ---------------------------------------------------------------------------------------------
    ll = mdl.log_likelihood(beta, sig)
    by = 8.0 * n * m_local + 16.0 * n
    g.estep_set_rowstats(False)
    try:
        mdl.log_likelihood(beta, sig)
        t_stream = env.max(timed(ctx, lambda: mdl.log_likelihood(beta, sig), reps=5, flush=True))
    finally:
        g.estep_set_rowstats(True)
    ll_default = mdl.log_likelihood(beta, sig)
    t_def = env.max(timed(ctx, lambda: mdl.log_likelihood(beta, sig), reps=5))
    B64 = np.asfortranarray(beta[:, None] * (1 + 1e-6 * np.arange(64))[None, :])
    mdl.log_likelihood_batch(B64, np.full(64, sig))
    t_b = env.max(timed(ctx, lambda: mdl.log_likelihood_batch(B64, np.full(64, sig)), reps=3))
    out["estep"] = {"evals_per_s": 1e3 / t_def, "evals_per_s_batched": 64e3 / t_b, "evals_per_s_stream_cold": 1e3 / t_stream,
                    "default_path": "stream of the factor matrix exp(s zd)" if fam == "binomial" else "row statistics of zd, O(n) per evaluation",
                    "roofline": {"bound": "hbm", "kernel": "loglik_logit_factor_kernel" if fam == "binomial" else "loglik_kernel<%d> (streaming)" % fl,
                                 "achieved": by / t_stream / 1e6, "peak": env.hbm, "unit": "GB/s", "frac": by / t_stream / 1e6 / env.hbm,
                                 "bytes_per_launch": by, "ms": t_stream, "note": "per GPU; L2 flushed before every launch"}}
    # ---- K3: MCNR step ----------------------------------------------------------------------------------------------------
    nr = mdl.mcnr(beta, sig)
    byn = 8.0 * n * m_local + 8.0 * n * (P + 2)
    t_nr = env.max(timed(ctx, lambda: mdl.mcnr(beta, sig), reps=3, flush=True))
    out["mcnr"] = {"steps_per_s": 1e3 / t_nr,
                   "roofline": {"bound": "hbm", "kernel": "mcnr_tma_kernel + mcnr_tail_kernel (whole mcnr() call incl. x'beta and the read-back)", "achieved": byn / t_nr / 1e6,
                                "peak": env.hbm, "unit": "GB/s", "frac": byn / t_nr / 1e6 / env.hbm, "bytes_per_launch": byn, "ms": t_nr}}
    # ---- K4 + K5: mvn_ll at a new theta (factorisation not cached) --------------------------------------------------------
    th1 = theta * (1 + 1e-4)
    dl = cv.loglik_model(th1, mdl)
    k = [1]

    def mvn_new():
        k[0] += 1
        cv.loglik_model(theta * (1 + 1e-4 * k[0]), mdl)

    def fac_new():
        k[0] += 1
        cv.logdet(theta * (1 + 1e-4 * k[0]))
    t_d = env.max(timed(ctx, mvn_new, reps=3))
    t_f = env.max(timed(ctx, fac_new, reps=3))
    nb = np.asarray(cfg["cov"]).reshape(-1, 5)
    nb = nb[np.unique(nb[:, 0], return_index=True)[1], 1].astype(float)
    fl_fac = float(np.sum(nb ** 3) / 3)
    fl_solve = float(np.sum(nb ** 2) * m_local)
    out["mvn_ll"] = {"evals_per_s": 1e3 / t_d, "ms": t_d, "factor_ms": t_f, "max_block": int(nb.max())}
    if large_block:
        # the streaming evaluation (forward substitution of all m sample columns) is the roofline kernel; the default on one rank with m >= 2 n_b is
        # the Cholesky factor of the block's Gram matrix (built once per sample matrix): n_b^3 / 3 flop per theta instead of n_b^2 m
        gram_used = env.world == 1 and m_local >= 2 * int(nb.max())
        t_mvn_stream = t_d
        if gram_used:
            g.cov_set_gram(False)
            try:
                cv.loglik_model(th1, mdl)
                t_mvn_stream = env.max(timed(ctx, mvn_new, reps=3))
            finally:
                g.cov_set_gram(True)
            out["mvn_ll"].update({"default_path": "Cholesky factor of the block's Gram matrix (one SYRK + factorisation per sample matrix), triangular right-hand sides per theta",
                                  "ms_streaming": t_mvn_stream, "executed_flops_default": fl_fac + float(np.sum(nb ** 3) / 3)})
        out["mvn_ll"]["roofline"] = {"bound": "fp64", "kernel": "blocked Cholesky + blocked forward substitution of the m sample columns (DMMA)" + (" — the streaming evaluation, Gram path off" if gram_used else ""),
                                     "achieved": (fl_fac + fl_solve) / t_mvn_stream / 1e9,
                                     "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": (fl_fac + fl_solve) / t_mvn_stream / 1e9 / FP64_PEAK_TFLOPS,
                                     "flops_per_launch": fl_fac + fl_solve, "ms": t_mvn_stream, "factor_tflops": fl_fac / t_f / 1e9, "factor_frac": fl_fac / t_f / 1e9 / FP64_PEAK_TFLOPS,
                                     "solve_tflops": fl_solve / max(t_mvn_stream - t_f, 1e-6) / 1e9}
    else:
        # default path: Gram matrices of the samples, independent of m; the streaming forward substitution is the HBM-bound kernel
        g.cov_set_gram(False)
        try:
            cv.loglik_model(th1, mdl)
            t_s = env.max(timed(ctx, mvn_new, reps=3, flush=True))
        finally:
            g.cov_set_gram(True)
        byu = 8.0 * Q * m_local
        T8 = np.asfortranarray(theta[:, None] * (1 + 1e-4 * np.arange(1, 65))[None, :])
        cv.loglik_model_batch(T8, mdl)
        t_db = env.max(timed(ctx, lambda: cv.loglik_model_batch(T8, mdl), reps=3))
        out["mvn_ll"].update({"default_path": "Gram matrices of the samples (one pass per sample matrix), evaluation independent of m",
                              "evals_per_s_batched": 64e3 / t_db,
                              "roofline": {"bound": "hbm", "kernel": "quad_small_kernel (streaming forward substitution, Gram path off)", "achieved": byu / t_s / 1e6,
                                           "peak": env.hbm, "unit": "GB/s", "frac": byu / t_s / 1e6 / env.hbm, "bytes_per_launch": byu, "ms": t_s}})

    # ---- parity of the E-step pieces against the oracle on the same columns -----------------------------------------------
    if check:
        t0 = time.perf_counter()
        xb = cfg["X"] @ beta
        ll_sum = 0.0; w = np.zeros(n); wu = np.zeros(n); sg = 0.0; mv_sum = 0.0
        Lh = None
        if large_block:
            import scipy.linalg as sla
            D1 = dense_D(cfg, th1)
            Lh = np.linalg.cholesky(D1)
        for a, b in col_chunks(m_local, max(n, Q)):
            Uc = np.asfortranarray(U[:, a:b])
            zd = z_apply(cfg, Uc)
            _, ps = oracle.loglik_zd(zd, xb, cfg["y"], sig, fl, per_sample=True)
            ll_sum += float(np.sum(ps))
            w_c, wu_c, sg_c = oracle.mcnr_sums_zd(zd, xb, cfg["y"], sig, fl)
            w += w_c; wu += wu_c; sg += sg_c
            if large_block:
                Wc = sla.solve_triangular(Lh, Uc, lower=True, check_finite=False)
                mv_sum += float(-0.5 * np.sum(Wc * Wc))
            else:
                mv_sum += oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], th1, Uc) * (b - a)
        ll_ref = env.sum(ll_sum) / m_total
        w = env.sum(w); wu = env.sum(wu); sg = env.sum(sg)
        X = cfg["X"]
        xtwx_ref = X.T @ (w[:, None] * X) / m_total
        score_ref = X.T @ wu / m_total
        mv_ref = env.sum(mv_sum) / m_total
        if large_block:
            mv_ref += -0.5 * Q * np.log(2 * np.pi) - float(np.sum(np.log(np.diag(Lh))))
        parity["loglik"] = {"gpu": ll, "oracle": ll_ref, "rel_err": abs(ll - ll_ref) / abs(ll_ref), "rel_err_default_path": abs(ll_default - ll_ref) / abs(ll_ref), "tol": 1e-10}
        parity["mcnr"] = {"xtwx_rel_err": rel(nr["xtwx"], xtwx_ref), "score_rel_err": float(np.max(np.abs(nr["score"] - score_ref)) / max(np.max(np.abs(score_ref)), np.max(np.abs(xtwx_ref)) * 1e-3)),
                          "sigma_rel_err": abs(nr["sigma"] - sg / m_total) / (sg / m_total), "tol": 1e-10}
        parity["mvn_ll"] = {"gpu": dl, "oracle": mv_ref, "rel_err": abs(dl - mv_ref) / abs(mv_ref), "tol": 1e-10,
                            "checker": "LAPACK Cholesky + triangular solve on the oracle's formula (mcmldmatrix.h:67-75)" if large_block else "oracle.mvn_loglik"}
        if large_block:
            Lg = cv.genD(th1, chol=True)
            import scipy.linalg.lapack as lap
            anorm = float(np.max(np.sum(np.abs(D1), axis=0)))
            rcond, info = lap.dpocon(Lh, anorm, uplo='L')
            kappa = 1.0 / rcond if rcond > 0 else float("inf")
            err = float(np.max(np.abs(Lg - Lh)) / np.max(np.abs(Lh)))
            parity["chol"] = {"max_rel_err": err, "kappa_1": kappa, "err_over_kappa_eps": err / (kappa * 2.220446049250313e-16), "tol": "1e-12 kappa",
                              "ok": bool(err <= 1e-12 * max(kappa, 1.0))}
            del Lg, D1
        parity["oracle_s"] = round(time.perf_counter() - t0, 2)

    # ---- fp32 mode: the same E-step on float storage of zd / F (half the bytes), against the same fp64 oracle values at 1e-5 -------------
    if fam in ("binomial", "poisson", "gaussian"):
        m32 = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], fam, link, precision="fp32")
        m32.set_u(U, m_total=m_total)
        ll32 = m32.log_likelihood(beta, sig)
        nr32 = m32.mcnr(beta, sig)
        by32 = 4.0 * n * m_local + 8.0 * n
        g.estep_set_rowstats(False)
        try:
            m32.log_likelihood(beta, sig)
            t32 = env.max(timed(ctx, lambda: m32.log_likelihood(beta, sig), reps=5, flush=True))
        finally:
            g.estep_set_rowstats(True)
        t32n = env.max(timed(ctx, lambda: m32.mcnr(beta, sig), reps=3, flush=True))
        out["fp32"] = {"dtype": "f32 storage of zd / F, f64 accumulation (gmb_model_create_prec(..., 32))",
                       "estep_evals_per_s_stream_cold": 1e3 / t32, "speedup_vs_f64_stream": t_stream / t32,
                       "mcnr_steps_per_s": 1e3 / t32n, "mcnr_speedup_vs_f64": t_nr / t32n,
                       "roofline": {"bound": "hbm", "achieved": by32 / t32 / 1e6, "peak": env.hbm, "unit": "GB/s", "frac": by32 / t32 / 1e6 / env.hbm,
                                    "bytes_per_launch": by32, "ms": t32},
                       "mcnr_roofline": {"bound": "hbm", "achieved": (by32 + 8.0 * n * P) / t32n / 1e6, "peak": env.hbm, "unit": "GB/s",
                                         "frac": (by32 + 8.0 * n * P) / t32n / 1e6 / env.hbm, "ms": t32n}}
        if check:
            parity["fp32"] = {"loglik_rel_err": abs(ll32 - ll_ref) / abs(ll_ref), "mcnr_xtwx_rel_err": rel(nr32["xtwx"], xtwx_ref),
                              "mcnr_sigma_rel_err": abs(nr32["sigma"] - sg / m_total) / (sg / m_total), "tol": 1e-5}
        m32.close()

    # ---- K6: sampler -------------------------------------------------------------------------------------------------------
    h = cfg["hmc"]
    C_local = max(1, cfg["chains_total"] // world)
    L = cfg["L"] if not large_block else None
    if L is None:
        L = cv.genD(theta, chol=True)
    mdl.hmc_sample(L, beta, sig, warmup=1, nsamp_per_chain=1, lam=0.05, max_steps=4, n_chains=C_local, chain_offset=rank * C_local, seed=11,
                   keep_on_device=True, want_u=False)          # uploads L, forms Z L and its sparse forms; warms the kernels up
    res = mdl.hmc_sample(None, beta, sig, warmup=h["warmup"], nsamp_per_chain=h["nsamp"], lam=h["lam"], max_steps=h["max_steps"],
                         target_accept=h["target_accept"], n_chains=C_local, chain_offset=rank * C_local, seed=12, keep_on_device=True, want_u=False)
    st = res["stats"]
    t_h = env.max(st["kernel_ms"])
    cols = C_local * (h["nsamp"] + 1)
    names = {1: "two fused-epilogue DMMA contractions per leapfrog step (K6')", 2: "on-chip dense (K6)", 3: "structure-aware sparse Z L (K6s / K6c)"}
    kern = names[st["kernel_variant"]] + (" with Z applied in sparse form (K6f)" if st["factored"] else "") + (", trajectories decomposed over %d component groups (K6c)" % st["component_groups"] if st["component_groups"] else "")
    lf = env.sum(st["leapfrog_total"])
    out["sampler"] = {"kernel": kern, "chains_total": C_local * world, "warmup": h["warmup"], "columns_per_chain": h["nsamp"] + 1,
                      "u_samples_per_s": env.sum(cols) / (t_h * 1e-3), "leapfrog_per_s": lf / (t_h * 1e-3), "ms": t_h,
                      "accept_rate": st["accept_rate"], "steps_mean": st["steps_mean"], "rows_used": st["rows_used"], "zl_nonzeros": st["zl_nonzeros"],
                      "algorithmic_tflops": lf * 4.0 * n * Q / t_h / 1e9,
                      "note": "R-default trajectory settings (lambda 5, <= 100 leapfrog steps, target 0.95) with a SHORTENED warm-up / draw count so the bench stays bounded: "
                              "u-samples/s is columns produced / kernel time at these counts, leapfrog steps/s is the rate that carries over to any proposal count"}
    if st["kernel_variant"] == 1 and st["factored"]:
        tri = 0.5 + 64.0 / Q                                          # L is triangular: the k tiles of its zero triangle are skipped
        ex = lf / world * 4.0 * Q * Q * tri / (st["kernel_ms"] * 1e-3) / 1e12
        out["sampler"]["roofline"] = {"bound": "fp64", "achieved": ex, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ex / FP64_PEAK_TFLOPS,
                                      "note": "executed flops per GPU: 4 Q^2 (tri) per leapfrog step and chain (W = L V', G = L' T)"}
    elif st["kernel_variant"] == 3:
        ex = lf / world * (4.0 * st["zl_nonzeros"] + 20.0 * st["rows_used"] + 8.0 * Q) / (st["kernel_ms"] * 1e-3) / 1e12
        out["sampler"]["roofline"] = {"bound": "fp64", "achieved": ex, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ex / FP64_PEAK_TFLOPS,
                                      "note": "executed FP64 flops per GPU (4 per non-zero of Z L + residual per row + update per column); latency / shared-memory bound, see DESIGN"}
    # E-step on the sampler's own device-resident columns (the MCML data flow: sample -> zd -> objective)
    t0 = time.perf_counter(); mdl.use_device_u(); ctx.sync(); out["sampler"]["zd_from_device_samples_s"] = env.max(time.perf_counter() - t0)
    ll_dev = mdl.log_likelihood(beta, sig)
    if check:
        t0 = time.perf_counter()
        Ud = mdl.get_u(0, cols)
        xb = cfg["X"] @ beta
        s = 0.0
        for a, b in col_chunks(cols, max(n, Q)):
            _, ps = oracle.loglik_zd(z_apply(cfg, np.asfortranarray(Ud[:, a:b])), xb, cfg["y"], sig, fl, per_sample=True)
            s += float(np.sum(ps))
        ll_dev_ref = env.sum(s) / env.sum(float(cols))
        parity["loglik_on_sampled_u"] = {"gpu": ll_dev, "oracle": ll_dev_ref, "rel_err": abs(ll_dev - ll_dev_ref) / abs(ll_dev_ref), "tol": 1e-10}
        # target density and gradient at random whitened states, and chain 0 of this rank state by state (short trajectories)
        ZL = np.asfortranarray(L[cfg["zidx"], :]) if cfg["zidx"] is not None else np.asfortranarray(cfg["Z"] @ L)
        rng = np.random.default_rng(5 + rank)
        V = np.asfortranarray(0.3 * rng.standard_normal((Q, 3)))
        lp, gr = mdl.log_prob_grad(None, beta, sig, V)
        lp_ref = np.array([oracle.log_prob(ZL, xb, cfg["y"], sig, fl, V[:, c]) for c in range(3)])
        gr_ref = np.stack([oracle.log_grad(ZL, xb, cfg["y"], sig, fl, V[:, c]) for c in range(3)], axis=1)
        parity["log_prob"] = {"rel_err": rel(lp, lp_ref), "tol": 1e-10}
        parity["log_grad"] = {"rel_err": rel(gr, gr_ref), "tol": 1e-10}
        wq, nq, lamq, msq = 2, 2, 0.03, 6
        rq = mdl.hmc_sample(None, beta, sig, warmup=wq, nsamp_per_chain=nq, lam=lamq, max_steps=msq, target_accept=0.9, n_chains=C_local,
                            chain_offset=rank * C_local, seed=4242, keep_on_device=False, want_u=False, want_v=True)
        ref = oracle.hmc_chain(ZL, L, xb, cfg["y"], sig, fl, wq, nq, lamq, msq, 0.9, 4242, chain=rank * C_local, want_u=False)
        sq = rq["stats"]
        parity["chain"] = {"max_abs_err": float(np.max(np.abs(rq["v"][:, :nq + 1] - ref["v"]))), "tol": 1e-7, "proposals": wq + nq, "chains_in_launch": C_local,
                           "kernel_variant": sq["kernel_variant"], "factored": sq["factored"], "component_groups": sq["component_groups"],
                           "same_kernel_as_timed_run": bool(sq["kernel_variant"] == st["kernel_variant"] and sq["factored"] == st["factored"]
                                                            and (sq["component_groups"] > 0) == (st["component_groups"] > 0))}
        parity["sampler_oracle_s"] = round(time.perf_counter() - t0, 2)
        ok = (parity["loglik"]["rel_err"] <= 1e-10 and parity["loglik"]["rel_err_default_path"] <= 1e-10 and parity["mcnr"]["xtwx_rel_err"] <= 1e-10
              and parity["mcnr"]["score_rel_err"] <= 1e-10 and parity["mcnr"]["sigma_rel_err"] <= 1e-10 and parity["mvn_ll"]["rel_err"] <= 1e-10
              and parity["loglik_on_sampled_u"]["rel_err"] <= 1e-10 and parity["log_prob"]["rel_err"] <= 1e-10 and parity["log_grad"]["rel_err"] <= 1e-10
              and parity["chain"]["max_abs_err"] <= 1e-7 and parity.get("chol", {"ok": True})["ok"]
              and max(parity.get("fp32", {"loglik_rel_err": 0.0})["loglik_rel_err"], parity.get("fp32", {"mcnr_xtwx_rel_err": 0.0})["mcnr_xtwx_rel_err"]) <= 1e-5)
        parity["ok"] = bool(ok)
    out["parity"] = parity if check else None
    mdl.close(); cv.close()
    return out


def dense_D(cfg, theta):
    """numpy D(theta) of a single dense fexp block (function id 13: theta_1 exp(-d / theta_2)) for the LAPACK checker."""
    cov = np.asarray(cfg["cov"]).reshape(-1, 5)
    assert cov.shape[0] == 1 and int(cov[0, 2]) == 13, "LAPACK checker: one fexp block expected"
    nloc = int(cov[0, 1])
    x = np.asarray(cfg["data"][:nloc]); y = np.asarray(cfg["data"][nloc:2 * nloc])
    d = np.sqrt((x[:, None] - x[None, :]) ** 2 + (y[:, None] - y[None, :]) ** 2)
    d *= -1.0 / theta[1]
    np.exp(d, out=d)
    d *= theta[0]
    return d


def main():
    import glmmrmcml_b200 as g
    size = "small" if "--small" in sys.argv else "full"
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["C3", "C4", "C5"]
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6650.0
    ctx = g.Context(0)
    env = Env(g, ctx, hbm_gbs=hbm)
    for nm in which:
        print(json.dumps(run_config(nm, env, size=size)), flush=True)


if __name__ == "__main__":
    main()
