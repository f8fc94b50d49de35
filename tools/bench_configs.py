#!/usr/bin/env python
"""Scale runs of the hot-path pieces on BASELINE.json's larger configurations (C3, C4, C5) — parity-test cases in
tests/, timed here at (near) full size against the roofline that bounds each kernel:

  zd = Z u build (K1)            FP64 DMMA GEMM, 2 n Q m flop
  E-step log-likelihood (K2)     HBM stream, 8 n m + 16 n bytes
  MCNR sufficient sums (K3)      HBM stream, 8 n m + 8 n (P + 2) bytes
  mvn_ll (K4 + K5)               D(theta) build + Cholesky + solves: sum_b n_b^3/3 + Q n_b m flop / 8 Q m bytes
  sampler (K6 / K6')             4 n Q flop per leapfrog step and chain

  python tools/bench_configs.py [C3 C4 C5] [--small]     -> one JSON line per configuration
The sample matrix is a tiling of a small seeded draw (the arithmetic cost of the kernels does not depend on the values)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth

FP64_PEAK = 37.1      # TFLOP/s, profiles/r01_microbench_fp64.txt
try:
    HBM_PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM_PEAK = 6650.0

small = "--small" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["C3", "C4", "C5"]
ctx = g.Context(0)


def tiled_u(cfg, m):
    U0 = cfg["U"]
    reps = (m + U0.shape[1] - 1) // U0.shape[1]
    return np.asfortranarray(np.tile(U0, (1, reps))[:, :m])


def run(name):
    t0 = time.perf_counter()
    if name == "C3":
        nloc = 1000 if small else 4000
        cfg = synth.config3(nloc=nloc, m=64); m = 250; chains, hw, hn = 256, 6, 2
    elif name == "C4":
        ncl = 200 if small else 1000
        cfg = synth.config4(ncl=ncl, nt=10, k=1, m=64); m = 2000 if small else 20000; chains, hw, hn = 1024, 6, 2
    else:
        nloc = 1000 if small else 5000
        cfg = synth.config5(nloc=nloc, nobs=10, m=64); m = 1000 if small else 10000; chains, hw, hn = 1024, 6, 2
    n, P, Q = cfg["n"], cfg["P"], cfg["Q"]
    U = tiled_u(cfg, m)
    t_gen = time.perf_counter() - t0
    out = {"config": name, "n": n, "P": P, "Q": Q, "m": m, "family": cfg["family"], "host_setup_s": round(t_gen, 2)}
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    cv = g.Covariance(ctx, cfg["cov"], cfg["data"], cfg["eff_range"])
    # K1: zd = Z u  (set_u = H2D of U + GEMM; the GEMM alone is timed by re-using the device copy)
    t0 = time.perf_counter(); mdl.set_u(U); out["set_u_s"] = time.perf_counter() - t0
    out["zd_bytes"] = 8.0 * n * m
    # K2
    beta = cfg["beta"]; sig = cfg.get("sigma", 1.0)
    g.estep_set_rowstats(False)                      # the streaming kernel (the default path of poisson / gaussian is O(n), timed below)
    mdl.log_likelihood(beta, sig)
    ts = []
    for r in range(5):
        ctx.flush_l2(); ctx.sync()
        ctx.timer_start(); mdl.log_likelihood(beta * (1 + 1e-6 * r), sig); ts.append(ctx.timer_stop())
    t_ll = float(np.median(ts)); by = 8.0 * n * m + 16.0 * n
    g.estep_set_rowstats(True)
    mdl.log_likelihood(beta, sig); td = []
    for r in range(5):
        ctx.sync(); ctx.timer_start(); mdl.log_likelihood(beta * (1 + 1e-6 * r), sig); td.append(ctx.timer_stop())
    t_def = float(np.median(td))
    out["loglik"] = {"stream_ms": t_ll, "stream_evals_per_s": 1e3 / t_ll, "GBps": by / t_ll / 1e6, "frac_hbm": by / t_ll / 1e6 / HBM_PEAK,
                     "default_path_ms": t_def, "default_path_evals_per_s": 1e3 / t_def,
                     "default_path": "factor matrix stream" if cfg["family"] == "binomial" else "row statistics, O(n) per evaluation"}
    # batched evaluations (what the optimiser and the Hessian stencil issue): 64 parameter vectors per call
    B64 = np.asfortranarray(beta[:, None] * (1 + 1e-6 * np.arange(64))[None, :])
    mdl.log_likelihood_batch(B64, np.full(64, sig)); ctx.sync()
    ctx.timer_start(); mdl.log_likelihood_batch(B64, np.full(64, sig)); t_b = ctx.timer_stop()
    out["loglik"]["batched_evals_per_s"] = 64e3 / t_b
    # K3
    mdl.mcnr(beta, sig); ts = []
    for r in range(3):
        ctx.flush_l2(); ctx.sync()
        ctx.timer_start(); mdl.mcnr(beta, sig); ts.append(ctx.timer_stop())
    t_nr = float(np.median(ts)); byn = 8.0 * n * m + 8.0 * n * (P + 2)
    out["mcnr"] = {"ms": t_nr, "steps_per_s": 1e3 / t_nr, "GBps": byn / t_nr / 1e6, "frac_hbm": byn / t_nr / 1e6 / HBM_PEAK}
    # K4 + K5: every evaluation at a new theta (factorisation not cached)
    th = cfg["theta"]; cv.loglik_model(th, mdl); ts = []
    for r in range(3):
        ctx.timer_start(); cv.loglik_model(th * (1 + 1e-4 * (r + 1)), mdl); ts.append(ctx.timer_stop())
    t_d = float(np.median(ts))
    tf = []
    for r in range(3):
        ctx.sync(); ctx.timer_start(); cv.logdet(th * (1 + 1e-4 * (r + 5))); tf.append(ctx.timer_stop())
    t_f = float(np.median(tf))
    nb = cfg["cov"][:, 1][np.unique(cfg["cov"][:, 0], return_index=True)[1]].astype(float)
    fl_d = float(np.sum(nb ** 3) / 3 + np.sum(nb ** 2) * m)
    out["mvn_ll"] = {"ms": t_d, "evals_per_s": 1e3 / t_d, "blocks": int(nb.size), "max_block": int(nb.max()), "flop": fl_d,
                     "TFLOPs": fl_d / t_d / 1e9, "GBps_of_U": 8.0 * Q * m / t_d / 1e6,
                     "factor_ms": t_f, "factor_TFLOPs": float(np.sum(nb ** 3) / 3) / t_f / 1e9}
    # K6 / K6'
    L = cv.genD(th, chol=True)
    mdl.hmc_sample(L, beta, sig, warmup=2, nsamp_per_chain=1, lam=0.05, max_steps=10, n_chains=chains, seed=1, keep_on_device=True, want_u=False)
    res = mdl.hmc_sample(None, beta, sig, warmup=hw, nsamp_per_chain=hn, lam=0.05, max_steps=10, target_accept=0.9, n_chains=chains, seed=2,
                         keep_on_device=True, want_u=False)
    st = res["stats"]; fl_h = st["leapfrog_total"] * 4.0 * n * Q
    names = {1: "two-GEMM (K6')", 2: "on-chip dense (K6)", 3: "structure-aware (K6s)"}
    out["hmc"] = {"kernel": names[st["kernel_variant"]], "chains": chains, "ms": st["kernel_ms"], "rows_used": st["rows_used"], "zl_nonzeros": st["zl_nonzeros"],
                  "leapfrog_per_s": st["leapfrog_total"] / st["kernel_ms"] * 1e3,
                  "algorithmic_TFLOPs": fl_h / st["kernel_ms"] / 1e9, "accept": st["accept_rate"], "steps_mean": st["steps_mean"]}
    out["hmc"]["factored"] = st["factored"]; out["hmc"]["component_groups"] = st["component_groups"]
    if st["kernel_variant"] != 3 and st["factored"]:
        # Z sparse, L dense: the contractions are Q x Q; the dense n x Q contraction on the same model for reference
        # the Cholesky factor is triangular: the k tiles of its zero triangle are skipped (128-row tiles), so about half of 4 Q^2 is executed
        tri = 0.5 + 64.0 / Q
        out["hmc"]["executed_TFLOPs"] = st["leapfrog_total"] * 4.0 * Q * Q * tri / st["kernel_ms"] / 1e9
        out["hmc"]["frac_fp64_executed"] = out["hmc"]["executed_TFLOPs"] / FP64_PEAK
        g.hmc_set_factored(False)
        try:
            rd = mdl.hmc_sample(None, beta, sig, warmup=1, nsamp_per_chain=1, lam=0.05, max_steps=10, target_accept=0.9, n_chains=chains, seed=2,
                                keep_on_device=True, want_u=False)["stats"]
        finally:
            g.hmc_set_factored(True)
        fd = rd["leapfrog_total"] * 4.0 * n * Q
        out["hmc"]["dense_kernel"] = {"kernel": names[rd["kernel_variant"]], "ms": rd["kernel_ms"], "leapfrog_per_s": rd["leapfrog_total"] / rd["kernel_ms"] * 1e3,
                                      "TFLOPs": fd / rd["kernel_ms"] / 1e9, "frac_fp64": fd / rd["kernel_ms"] / 1e9 / FP64_PEAK}
        res = mdl.hmc_sample(None, beta, sig, warmup=hw, nsamp_per_chain=hn, lam=0.05, max_steps=10, target_accept=0.9, n_chains=chains, seed=2,
                             keep_on_device=True, want_u=False)
    elif st["kernel_variant"] != 3:
        out["hmc"]["frac_fp64"] = fl_h / st["kernel_ms"] / 1e9 / FP64_PEAK
    else:
        # the dense kernel on the same model, for reference (fewer proposals: it is orders of magnitude slower here)
        g.hmc_set_variant(1)
        try:
            rd = mdl.hmc_sample(None, beta, sig, warmup=1, nsamp_per_chain=1, lam=0.05, max_steps=10, target_accept=0.9, n_chains=chains, seed=2,
                                keep_on_device=True, want_u=False)["stats"]
        finally:
            g.hmc_set_variant(0)
        fd = rd["leapfrog_total"] * 4.0 * n * Q
        out["hmc"]["dense_kernel"] = {"kernel": names[rd["kernel_variant"]], "ms": rd["kernel_ms"], "leapfrog_per_s": rd["leapfrog_total"] / rd["kernel_ms"] * 1e3,
                                      "TFLOPs": fd / rd["kernel_ms"] / 1e9, "frac_fp64": fd / rd["kernel_ms"] / 1e9 / FP64_PEAK}
        res = mdl.hmc_sample(None, beta, sig, warmup=hw, nsamp_per_chain=hn, lam=0.05, max_steps=10, target_accept=0.9, n_chains=chains, seed=2,
                             keep_on_device=True, want_u=False)
    # K1: zd = Z u for the sampler's device-resident draws (chains * (hn + 1) columns)
    mh = chains * (hn + 1)
    ctx.sync(); ctx.timer_start(); mdl.use_device_u(); t = ctx.timer_stop()
    nnz_z = int(np.count_nonzero(cfg["Z"][: min(n, 2000)])) * (n / min(n, 2000))
    sparse_zd = Q >= 64 and nnz_z * 16 <= float(n) * Q                 # the library's criterion (gmb_zell_ensure): gather instead of a contraction
    out["zd_build"] = {"cols": mh, "ms": t, "path": "gather through the sparse form of Z" if sparse_zd else "DMMA contraction",
                       "GBps_written": 8.0 * n * mh / t / 1e6, "algorithmic_TFLOPs": 2.0 * n * Q * mh / t / 1e9}
    if not sparse_zd:
        out["zd_build"]["frac_fp64"] = 2.0 * n * Q * mh / t / 1e9 / FP64_PEAK
    mdl.close(); cv.close()
    print(json.dumps(out), flush=True)


for nm in which:
    run(nm)
