#!/usr/bin/env python
"""bench.py — headline benchmark of the glmmrMCML hot path (Monte-Carlo E-step + random-effect sampler) on B200.

Workload (BASELINE.json configs[1], "C2"): cluster RCT, binomial-logit, n = 500, P = 6, Q = 50 (10 blocks of
gr(cl)*ar1(t)), m = 10^4 Monte-Carlo samples per GPU, MCNR + Hessian standard errors.

One STEP = one pass of the hot path over one batch of m samples, the work of one MCML iteration with Hessian SEs:
  1. draw m samples of u with the batched HMC sampler (C chains x (warmup + m/C) proposals; R defaults of
     ModelMCML$mcmc_options: warmup 500, lambda 5, maxsteps 100, target_accept 0.95)            [mhmcmc.h:121-157]
  2. zd = Z u                                                                                     [mcmlmodel.h:286]
  3. one MCNR step (per-sample sufficient sums + Newton increment)                                [mcmloptim.h:198-236]
  4. N_D mvn_ll evaluations (the d_optim objective; D(theta) build + Cholesky + forward solves)   [mcmldmatrix.h:23-78]
  5. the 4k^2 = 256 point Hessian stencil: 256 E-step log-likelihood evaluations + 256 mvn_ll     [mcmloptim.h:333-355]
value = u-samples/s = (m x ranks) / step time; the E-step evaluations/s and sampler samples/s (ESS/s) of
BASELINE.json's metric are reported beside it, each timed in isolation in the same run.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
(N > 1: launched by torchrun, one rank per GPU; samples and chains are sharded over ranks — weak scaling.)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_PER_GPU = 10_000
N_CHAINS = 1000           # per GPU; 10 columns each (9 post-warm-up draws + the state after warm-up).  The sampler aggregates the 500
                          # observations into their 50 distinct rows of [X | Z] (aggregate.cu) and, Z L being sparse (150 non-zeros),
                          # runs the structure-aware kernel, one warp per chain (hmc_sparse.cu).  Chain count: a draw costs (warm-up +
                          # columns per chain) x ~95 leapfrog steps in sequence, 280 ns each up to one warp per scheduler (592), 350 ns
                          # with 1000 chains.  500 chains x 20 columns give more columns per second (13.7 ms per draw against 16.9 ms)
                          # but fewer effective samples per second (median ESS/s 74 k against 93 k: consecutive draws of a chain are
                          # correlated) — tools/hmc_chain_count_sweep.py; the E-step integrates over these columns, so ESS/s decides
HMC = dict(warmup=500, lam=5.0, max_steps=100, target_accept=0.95, adapt=100)
N_D_EVALS = 64            # mvn_ll evaluations of one d_optim (BOBYQA over 2 parameters takes 40-80)
N_HESS = 256              # 4 k^2, k = P + R = 8
D_BATCH = 5               # 2 R + 1: the central-difference stencil the optimiser evaluates as one batch (optim.cpp)
FP64_DMMA_PEAK_TFLOPS = 37.1   # measured on this pool's B200: profiles/r01_microbench_fp64.txt (tools/microbench_fp64.cu)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures (profiles/r02_ncu_hmc_hmc_lane.txt,
# profiles/r01_ncu_raw_extract_final.txt):
SAMPLER_DRAM_BYTES_PER_LAUNCH = {2: 446976.0,       # hmc_fused_kernel<3,13,4>: Z L, xb, y are read once, the chain runs out of shared memory
                                 3: 80384.0,         # hmc_sparse_kernel<3,32,2,5>: ELL arrays, xb and row weights are read once; the samples stay in L2
                                 "lane": 128768.0}   # hmc_lane_kernel<3,1,5,5,1>: block of Z L, xb and row weights read once; state in registers, samples stay in L2
# FP64 operations per leapfrog step and chain that the structure-aware kernel executes (FMA = 2): 4 per non-zero of Z L (eta and gradient),
# per row the residual (table exp 18 + Newton reciprocal 9 + 2 for binomial-logit; 20 poisson; 2 gaussian), per column 8 (gradient, leapfrog)
SPARSE_ROW_FLOPS = {"binomial": 29.0, "poisson": 20.0, "gaussian": 2.0}
LOGLIK_DRAM_BYTES_PER_LAUNCH = 1.000243e9 + 6.959e6  # loglik_logit_factor_kernel<double> on 1 GB of F (algorithmic bytes: 1.000008e9); profiles/r02_ncu_estep_loglik_logit.txt
MCNR_DRAM_BYTES_PER_LAUNCH = 1.000088e9 + 14.16e6     # mcnr_tma_kernel<3, double> on the same matrix (same capture)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def sampler_roofline(st, exec_tflops, algo_tflops, exec_per_step, cfg, Q, dense_probe, many_probe=None):
    """roofline object of the sampler launch of the timed step (the step's dominant kernel)."""
    kv = st["kernel_variant"]
    lane = kv == 3 and st.get("lane_components", 0) > 0
    common = {"achieved": exec_tflops, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": exec_tflops / FP64_DMMA_PEAK_TFLOPS,
              "algorithmic_tflops": algo_tflops, "frac_algorithmic": algo_tflops / FP64_DMMA_PEAK_TFLOPS,
              "traffic": SAMPLER_DRAM_BYTES_PER_LAUNCH.get("lane" if lane else kv), "flops_per_leapfrog_per_chain": 4.0 * cfg["n"] * Q,
              "executed_flops_per_leapfrog_per_chain": exec_per_step, "rows_used": st["rows_used"], "zl_nonzeros": st["zl_nonzeros"],
              "peak_source": "FP64 pipe peak measured on this pool's B200 (profiles/r01_microbench_fp64.txt: DMMA 37.1, DFMA 36.8 TFLOP/s, the same "
                             "pipe); MEASURED_PEAKS.json has no fp64 entry (the bf16 tensor peak does not apply to an fp64 kernel)"}
    if lane:
        common.update({
            "kernel": "hmc_lane_kernel<%s> (structure-aware sampler, lane per connected component: Z L is block diagonal in the aggregated rows, "
                      "%d blocks per chain; one lane integrates a whole block — rows, columns, momentum, position, ELL entries in registers — "
                      "with no exchange inside a trajectory; a warp-level sum per proposal for the accept test)" % (cfg["family"], st["lane_components"]),
            "bound": "fp64",
            "note": "achieved = FP64 flops the kernel EXECUTES per leapfrog step and chain (4 per non-zero of Z L + residual per distinct row + "
                    "leapfrog update per column; FMA = 2) / kernel time, against the FP64 pipe peak.  SURVEY 8d's algorithmic figure for the dense "
                    "contraction (4 n Q = 1e5 per step and chain) is `algorithmic_tflops`: the kernel does the same arithmetic on the 150 non-zeros "
                    "of the 50 distinct rows instead of 500 x 50 entries, so that figure exceeds the pipe peak.  1000 chains x 10 blocks = 313 warps "
                    "on 592 schedulers: every warp has a scheduler to itself and runs one dependent FP64 chain (ncu, profiles/r02_ncu_hmc_hmc_lane.txt: "
                    "FP64 pipe 32 % busy averaged over the SMs — 334 one-warp CTAs, 2.3 per SM —, 3.5 % of the warp slots) — bound by FP64 latency x trajectory length, not by pipe "
                    "throughput; `dense_kernel` is the DMMA kernel a model with a dense Z L runs",
            "dense_kernel": dense_probe})
        if many_probe and many_probe.get("kernel_variant") == 3:
            t = many_probe["leapfrog_per_s"] * exec_per_step / 1e12
            common["saturated"] = {"chains": many_probe["chains"], "ms": many_probe["ms"], "leapfrog_per_s": many_probe["leapfrog_per_s"],
                                   "achieved": t, "frac": t / FP64_DMMA_PEAK_TFLOPS,
                                   "note": "the same kernel with 4000 chains (two warps per scheduler, 104 proposals): its throughput regime; not the timed step"}
    elif kv == 3:
        common.update({
            "kernel": "hmc_sparse_kernel<%s> (structure-aware sampler: Z L in ELL form, one warp per chain, state and ELL entries in registers)" % cfg["family"],
            "bound": "fp64",
            "note": "achieved = FP64 flops the kernel EXECUTES per leapfrog step and chain (4 per non-zero of Z L + residual per distinct row + "
                    "leapfrog update per column; FMA = 2) / kernel time, against the FP64 pipe peak.  SURVEY 8d's algorithmic figure for the dense "
                    "contraction (4 n Q = 1e5 per step and chain) is `algorithmic_tflops`: the kernel does the same arithmetic on the 150 non-zeros "
                    "of the 50 distinct rows instead of 500 x 50 entries, so that figure exceeds the pipe peak.  With 1000 chains (1.7 warps per "
                    "scheduler) a leapfrog step is bound by its dependent chain (shared-memory exchange + exp + reciprocal: 550 cycles alone, ~700 shared), not by "
                    "pipe throughput; `dense_kernel` is the DMMA kernel a model with a dense Z L runs",
            "dense_kernel": dense_probe})
        if many_probe and many_probe.get("kernel_variant") == 3:
            t = many_probe["leapfrog_per_s"] * exec_per_step / 1e12
            common["saturated"] = {"chains": many_probe["chains"], "ms": many_probe["ms"], "leapfrog_per_s": many_probe["leapfrog_per_s"],
                                   "achieved": t, "frac": t / FP64_DMMA_PEAK_TFLOPS,
                                   "note": "the same kernel with 4000 chains (several warps per scheduler, 104 proposals): bound by the shared-memory "
                                           "pipe (ncu: 69 % busy already at 1000 chains), not by latency; not the timed step"}
    else:
        common.update({
            "kernel": "hmc_fused_kernel<%s> (on-chip sampler: eta = ZL v and grad = ZL^T r(eta) as FP64 DMMA)" % cfg["family"] if kv == 2
                      else "dgemm_kernel<EpiResid/EpiLeapfrog> (two fused-epilogue DMMA GEMMs per leapfrog step)",
            "bound": "tensor",
            "note": "achieved = tensor flops the kernel executes (4 rows_used Q per leapfrog step and chain) / kernel time",
            "saturated": dense_probe})
    return common


def stencil_points(rng, P, R, beta, theta, npts, h=1e-5):
    """npts distinct parameter vectors around (beta, theta) like the optimhess stencil (every point evaluated, no memo)."""
    B = beta[:, None] + h * rng.integers(-2, 3, size=(P, npts))
    B[:, 0] = beta
    B += 1e-9 * np.arange(npts)[None, :]            # all distinct
    T = theta[:, None] + h * rng.integers(-2, 3, size=(R, npts)) + 1e-9 * np.arange(npts)[None, :]
    return np.asfortranarray(B), np.asfortranarray(T)


def ess_chains(x):
    """Effective sample size of C independent chains of N draws each (x: chains x draws) from the spread of the chain means:
    Var(chain mean) = sigma^2 tau / N, so ESS = C N / tau = C sigma^2 / Var(chain mean).  With hundreds of independent chains this
    is far less noisy than autocorrelation-based estimators on the few draws each chain contributes."""
    C, N = x.shape
    if C < 4:
        return float(C * N)
    s2 = x.var(ddof=1)                       # pooled variance (stationary marginal)
    vm = x.mean(axis=1).var(ddof=1)          # variance of the chain means
    if not (vm > 0 and s2 > 0):
        return float(C * N)
    return float(C * s2 / vm)


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(np.max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path (oracle FAITHFUL mode = the reference's loop structure) on a BOUNDED SAMPLE of the step,
# every piece executed for real — nothing is extrapolated into the arm's value
# ----------------------------------------------------------------------------------------------------------------------
CPU_SAMPLE_M = 2500       # Monte-Carlo samples of the bounded CPU sample (the GPU arm's step has 10^4)


def cpu_reference_sample(cfg, threads=None, m_s=CPU_SAMPLE_M, quick=False):
    """One pass of the C2 step at m_s samples on the host cores, as the reference runs it (src/mcml_full.cpp:83-126 with Hessian SEs):
    a single HMC chain of warmup + m_s proposals (mhmcmc.h:121-157), one MCNR step whose update_W(i) redoes the Z u GEMM for every sample
    (mcmloptim.h:213), N_D_EVALS + N_HESS mvn_ll evaluations that re-factorise every block per sample (mcmldmatrix.h:33-36) and N_HESS
    log-likelihood evaluations that redo the Z u GEMM (mcmlmodel.h:286).  Returns (seconds, detail); value = m_s / seconds."""
    import oracle
    oracle.build()
    if threads:
        oracle.set_threads(threads)
    nthr = oracle.max_threads() if not threads else threads
    if quick:
        m_s = min(m_s, 300)
    fl = oracle.flink(cfg["family"], cfg["link"])
    X, Z, y, beta, theta = cfg["X"], cfg["Z"], cfg["y"], cfg["beta"], cfg["theta"]
    U = np.asfortranarray(cfg["U"][:, :m_s])
    ZL = Z @ cfg["L"]
    xb = X @ beta
    P, R = cfg["P"], theta.size
    rng = np.random.default_rng(99)
    Bst, Tst = stencil_points(rng, P, R, beta, theta, N_HESS)
    _, Td = stencil_points(rng, P, R, beta, theta, N_D_EVALS, h=1e-3)
    d = {"m_sample": m_s, "threads": nthr}
    warm = 100 if quick else HMC["warmup"]
    t0 = time.perf_counter()
    ch = oracle.hmc_chain(ZL, cfg["L"], xb, y, 1.0, fl, warm, m_s, HMC["lam"], HMC["max_steps"], HMC["target_accept"], 12345, want_u=False)
    d["hmc_s"] = time.perf_counter() - t0
    d["hmc_proposals"] = warm + m_s
    d["hmc_steps_per_proposal"] = ch["total_steps"] / (warm + m_s)
    t0 = time.perf_counter()
    oracle.mcnr(X, Z, U, y, beta, 1.0, fl, faithful=True)
    d["mcnr_s"] = time.perf_counter() - t0
    n_ll = 16 if quick else N_HESS
    t0 = time.perf_counter()
    for k in range(n_ll):
        oracle.loglik_faithful(X, Z, U, y, Bst[:, k], 1.0, fl)
    d["estep_s"] = (time.perf_counter() - t0) * (N_HESS / n_ll)
    n_d = 16 if quick else N_D_EVALS + N_HESS
    Tall = np.concatenate([Td, Tst], axis=1)
    t0 = time.perf_counter()
    for k in range(n_d):
        oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], Tall[:, k], U, faithful=True)
    d["mvn_s"] = (time.perf_counter() - t0) * ((N_D_EVALS + N_HESS) / n_d)
    total = d["hmc_s"] + d["mcnr_s"] + d["estep_s"] + d["mvn_s"]
    return float(total), d


def cpu_hoisted_and_full_m(cfg, det, threads=None, quick=False):
    """Beside the arm's value: (i) the same sample with the reference's redundant work hoisted (zd built once, one factorisation per theta)
    — a stronger CPU baseline; (ii) what the faithful step costs at the GPU arm's full m = 10^4, from the measured pieces: sampler and
    evaluations scale linearly in m, MCNR with the exponent measured here on two sample sizes (labelled an estimate, not the arm's value)."""
    import oracle
    fl = oracle.flink(cfg["family"], cfg["link"])
    X, Z, y, beta, theta = cfg["X"], cfg["Z"], cfg["y"], cfg["beta"], cfg["theta"]
    m_s = det["m_sample"]
    U = np.asfortranarray(cfg["U"][:, :m_s])
    xb = X @ beta
    zd = oracle.gemm(Z, U)
    t0 = time.perf_counter(); oracle.loglik_zd(zd, xb, y, 1.0, fl); t_llh = time.perf_counter() - t0
    t0 = time.perf_counter(); oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], theta, U, faithful=False); t_dh = time.perf_counter() - t0
    t0 = time.perf_counter(); oracle.mcnr(X, Z, U, y, beta, 1.0, fl, faithful=False); t_mh = time.perf_counter() - t0
    hoisted = det["hmc_s"] + t_mh + N_HESS * t_llh + (N_D_EVALS + N_HESS) * t_dh
    m_half = max(64, m_s // 2)
    t0 = time.perf_counter(); oracle.mcnr(X, Z, np.asfortranarray(U[:, :m_half]), y, beta, 1.0, fl, faithful=True); t_half = time.perf_counter() - t0
    expo = float(np.log(det["mcnr_s"] / t_half) / np.log(m_s / m_half))
    f = M_PER_GPU / m_s
    per_prop = det["hmc_s"] / det["hmc_proposals"]
    full = per_prop * (HMC["warmup"] + M_PER_GPU) + det["mcnr_s"] * f ** expo + (det["estep_s"] + det["mvn_s"]) * f
    return {"hoisted_step_s": float(hoisted), "hoisted_value": m_s / hoisted, "hoisted_loglik_s_per_eval": t_llh,
            "mcnr_scaling_exponent_measured": expo, "reference_headers": reference_headers_timing(det, threads),
            "full_m_estimate": {"m": M_PER_GPU, "step_s": float(full), "value": M_PER_GPU / full,
                                "how": "sampler, log-likelihood and mvn_ll pieces x m / m_sample; MCNR x (m / m_sample)^exponent with the exponent measured in this run"}}


def reference_headers_timing(det, threads):
    """Beside the arm's value: the reference's OWN headers (oracle/_ref/libref_omp.so, compiled against the stand-in Eigen / Rcpp of
    oracle/shim) timed piece by piece next to the port on one small sample (tools/ref_headers_timing.py, ~5 s, in a subprocess: the OpenMP
    build of mcmloptim::mcnr races, so a fault there must not take the bench line with it).  `value_scaled` = the arm's own pieces multiplied by
    the measured headers / port ratios — what the arm would report if it timed the headers instead of the port."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_headers_timing.py"), str(int(threads or 1))],
                           capture_output=True, text=True, timeout=240)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": "rc %d: %s" % (r.returncode, (r.stderr or r.stdout)[-200:])}
        out = json.loads(lines[-1])
        if "unavailable" not in out:
            scaled = (det["hmc_s"] * out["hmc_s_per_proposal"]["headers_over_port"] + det["mcnr_s"] * out["mcnr_s"]["headers_over_port"]
                      + det["estep_s"] * out["loglik_s_per_eval"]["headers_over_port"] + det["mvn_s"] * out["mvn_ll_s_per_eval"]["headers_over_port"])
            out["step_s_scaled"] = float(scaled)
            out["value_scaled"] = det["m_sample"] / scaled
        return out
    except Exception as e:      # the port's number stands on its own
        return {"unavailable": str(e)[:200]}


def cpu_arm(cfg, threads, quick, m_s=CPU_SAMPLE_M):
    """(seconds per bounded sample, detail, kind, sample description) of the CPU arm."""
    t, det = cpu_reference_sample(cfg, threads=threads, m_s=m_s, quick=quick)
    sample = ("oracle FAITHFUL (the reference's loop structure incl. its redundant Z u GEMMs and per-sample Cholesky) on m_sample = %d of the step's "
              "10^4 samples, every piece run for real: one HMC chain of %d proposals, 1 MCNR step, %d log-likelihood and %d mvn_ll evaluations; "
              "value = m_sample / time (the reference's MCNR cost grows faster than m, so the smaller sample FAVOURS the CPU arm)"
              % (det["m_sample"], det["hmc_proposals"], N_HESS, N_D_EVALS + N_HESS))
    return t, det, "port", sample


def check_step_against_oracle(cfg, mdl, out, Bst, Tst, m_per_gpu, world, sum_over_ranks, n_pts=4):
    """The outputs of the timed step (MCNR sums, the first Hessian-stencil log-likelihoods and mvn_ll values) against the oracle on the
    sample columns the step drew.  Every rank evaluates the oracle on ITS columns; the sums are added over ranks like the device sums."""
    import oracle
    oracle.build()
    nr, ll, dl = out
    fl = oracle.flink(cfg["family"], cfg["link"])
    U = mdl.get_u(0, m_per_gpu)
    zd = oracle.gemm(cfg["Z"], U)
    m_tot = m_per_gpu * world
    X = cfg["X"]
    w, wu, sg = oracle.mcnr_sums_zd(zd, X @ cfg["beta"], cfg["y"], 1.0, fl)
    w = sum_over_ranks(w); wu = sum_over_ranks(wu); sg = sum_over_ranks(float(sg))
    xtwx = X.T @ (w[:, None] * X) / m_tot
    score = X.T @ wu / m_tot
    res = {"mcnr_xtwx_rel_err": float(np.max(np.abs(nr["xtwx"] - xtwx)) / np.max(np.abs(xtwx))),
           "mcnr_score_abs_err_over_xtwx": float(np.max(np.abs(nr["score"] - score)) / np.max(np.abs(xtwx))),
           "mcnr_sigma_rel_err": float(abs(nr["sigma"] - sg / m_tot) / (sg / m_tot))}
    e_ll, e_d = 0.0, 0.0
    for k in range(n_pts):
        _, ps = oracle.loglik_zd(zd, X @ Bst[:, k], cfg["y"], 1.0, fl, per_sample=True)
        ref = sum_over_ranks(float(np.sum(ps))) / m_tot
        e_ll = max(e_ll, abs(ll[k] - ref) / abs(ref))
        ref_d = sum_over_ranks(oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], Tst[:, k], U) * m_per_gpu) / m_tot
        e_d = max(e_d, abs(dl[k] - ref_d) / abs(ref_d))
    res.update(loglik_rel_err=float(e_ll), mvn_ll_rel_err=float(e_d), points=n_pts, tol=1e-10, ranks=world)
    res["ok"] = bool(max(res["mcnr_xtwx_rel_err"], res["mcnr_score_abs_err_over_xtwx"], res["mcnr_sigma_rel_err"], e_ll, e_d) <= 1e-10)
    return res


def vendor_comparators():
    """cuBLAS Dgemm / cuSOLVER Dpotrf / cuBLAS Dtrsm on the same GPU (tools/vendor_fp64, a stand-alone binary: tools only, the product
    library never links or calls the vendor BLAS).  None when the binary is missing."""
    exe = os.path.join(ROOT, "tools", "vendor_fp64")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-300:]}
    except Exception as e:
        return {"error": str(e)[:300]}


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller CPU samples (for tests)")
    ap.add_argument("--no-configs", action="store_true", help="skip the large configurations C3 / C4 / C5")
    ap.add_argument("--configs-size", default="full", choices=["full", "small"], help="size of the large configurations (small: for tests)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its version there) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    from glmmrmcml_b200 import synth
    cfg = synth.config2(m=M_PER_GPU)
    workload = {"workload": "C2: cluster RCT binomial-logit n=500 P=6 Q=50 (10 blocks gr(cl)*ar1(t)), m=10^4 samples per GPU, "
                            "MCNR + Hessian SEs",
                "m_per_gpu": M_PER_GPU, "chains_per_gpu": N_CHAINS, "hmc": HMC,
                "per_step": {"hmc_samples": M_PER_GPU, "mcnr_steps": 1, "loglik_evals": N_HESS, "mvn_ll_evals": N_D_EVALS + N_HESS},
                "l2": "every step rewrites all sampler state, U (4 MB) and zd (40 MB); the roofline probe streams 1 GB > L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        import oracle
        oracle.build()
        vals, det = [], None
        ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        # bounded sample: the whole --steps K run should end within a few minutes.  A step at m_sample = 2500 takes ~12 s on 16 cores; with more than
        # 12 timed steps the sample shrinks in proportion (never below 500).  value = m_sample / time either way — and a smaller sample favours the
        # CPU arm, whose MCNR cost grows faster than m.
        m_s = CPU_SAMPLE_M if args.steps <= 12 else max(500, int(CPU_SAMPLE_M * 12 / args.steps) // 100 * 100)
        for it in range(args.warmup + args.steps):
            # warm-up steps run a small sample (page-in, thread pool start); torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly
            t, d, kind, smp = cpu_arm(cfg, ncores, args.quick or it < args.warmup, m_s)
            if it >= args.warmup:
                vals.append(t); det = d; sample = smp
        t_step = float(np.median(vals)) if vals else float("nan")
        v = det["m_sample"] / t_step
        if not args.quick:
            det.update(cpu_hoisted_and_full_m(cfg, det, ncores))
        line = {"metric": "u-samples/s through sampler + E-step", "value": v, "unit": "u-samples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": workload,
                "cpu_baseline": {"value": v, "unit": "u-samples/s", "cores": det["threads"], "kind": kind, "sample": sample, "detail": det},
                "e2e": {"value": v, "unit": "u-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # ---------------- our arm ----------------
    import torch
    import glmmrmcml_b200 as g

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = g.Context(local_rank)
    if world > 1:
        ids = [g.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
    ctx.make_default()
    from tools import bench_configs as bc
    env = bc.Env(g, ctx, rank, world, dist, hbm_gbs=load_peaks()[0]["hbm_gbs"])      # rank / world and array reductions for the oracle checks

    P, Q, R = cfg["P"], cfg["Q"], cfg["theta"].size
    beta, theta = cfg["beta"], cfg["theta"]
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    cv = g.Covariance(ctx, cfg["cov"], cfg["data"], cfg["eff_range"])
    L = cv.genD(theta, chol=True)
    rng = np.random.default_rng(1234)
    Bst, Tst = stencil_points(rng, P, R, beta, theta, N_HESS)
    _, Td = stencil_points(rng, P, R, beta, theta, N_D_EVALS, h=1e-3)
    per = M_PER_GPU // N_CHAINS          # columns per chain incl. the post-warm-up state (k + 1)
    seed0 = 20221208

    first = {"done": False}

    def step(i):
        # 1. sampler (device-resident: L/ZL uploaded once)
        mdl.hmc_sample(None if first["done"] else L, beta, 1.0, warmup=HMC["warmup"], nsamp_per_chain=per - 1, lam=HMC["lam"],
                       max_steps=HMC["max_steps"], target_accept=HMC["target_accept"], adapt=HMC["adapt"], n_chains=N_CHAINS,
                       chain_offset=rank * N_CHAINS, seed=seed0 + i, keep_on_device=True, want_u=False)
        first["done"] = True
        mdl.use_device_u()                                   # 2. zd = Z u
        nr = mdl.mcnr(beta, 1.0)                             # 3.
        for k in range(0, N_D_EVALS, D_BATCH):               # 4. theta update: batches of 2R + 1 stencil points, as the optimiser issues them
            cv.loglik_model_batch(Td[:, k:k + D_BATCH], mdl)
        ll = mdl.log_likelihood_batch(Bst, np.ones(N_HESS))  # 5. Hessian stencil: one batch each
        dl = cv.loglik_model_batch(Tst, mdl)
        return nr, ll, dl

    for i in range(args.warmup):
        step(i)
    clocks = ClockSampler(local_rank)
    ctx.sync(); torch.cuda.synchronize(); barrier()
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count
    ctx.timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = step(args.warmup + i)
    ms = ctx.timer_stop()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = sum_over_ranks(ctx.launch_count - l0)
    ms_step = max_over_ranks(ms) / args.steps
    value = M_PER_GPU * world / (ms_step * 1e-3)
    # the timed step's own outputs against the oracle on the samples it drew (all ranks: local oracle sums, summed like the device sums)
    step_parity = check_step_against_oracle(cfg, mdl, out, Bst, Tst, M_PER_GPU, world, env.sum)
    assert step_parity["ok"], step_parity

    # ---- end to end through the reference-named C-ABI entry points with host buffers (what the R loop calls) ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    Zp = np.asfortranarray(pin(cfg["Z"].T).T); Xp = np.asfortranarray(pin(cfg["X"].T).T); yp = pin(cfg["y"]); Lp = np.asfortranarray(pin(L.T).T)
    start = np.concatenate([beta, theta, [1.0]])

    e2e_parts = {"mcmc_sample_s": [], "mcml_optim_s": [], "mcml_hess_s": []}

    def e2e_step(i):
        ta = time.perf_counter()
        u = g.mcmc_sample(Zp, Lp, Xp, yp, beta, cfg["family"], cfg["link"], HMC["warmup"], M_PER_GPU - 1, HMC["lam"], 1.0, 0, 500,
                          HMC["max_steps"], HMC["target_accept"], n_chains=N_CHAINS, seed=seed0 + 100 + i + 1000 * rank)
        tb = time.perf_counter()
        fit = g.mcml_optim(cfg["cov"], cfg["data"], cfg["eff_range"], Zp, Xp, yp, u, cfg["family"], cfg["link"], start, 0, True)
        tc = time.perf_counter()
        H = g.mcml_hess(cfg["cov"], cfg["data"], cfg["eff_range"], Zp, Xp, yp, u, cfg["family"], cfg["link"],
                        np.concatenate([fit["beta"], fit["theta"]]), 1e-5, 0)
        td = time.perf_counter()
        if i >= 0:
            e2e_parts["mcmc_sample_s"].append(tb - ta); e2e_parts["mcml_optim_s"].append(tc - tb); e2e_parts["mcml_hess_s"].append(td - tc)
        return u, fit, H

    n_e2e = max(2, min(args.steps, 3))
    e2e_step(-1)
    ctx.sync(); barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        u_last, fit_last, H_last = e2e_step(i)
    ctx.sync()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / n_e2e)
    e2e_val = M_PER_GPU * world / e2e_s
    base_b = (Zp.nbytes + Xp.nbytes + yp.nbytes)
    h2d = (base_b + Lp.nbytes) + 2 * (base_b + u_last.nbytes)
    d2h = u_last.nbytes + (P + R + 1) * 8 + (P + R) ** 2 * 8

    # ---- isolated components on rank 0's shard (device-resident), for BASELINE.json's two-part metric ----
    extra = {}
    res = mdl.hmc_sample(None, beta, 1.0, warmup=HMC["warmup"], nsamp_per_chain=per - 1, lam=HMC["lam"], max_steps=HMC["max_steps"],
                         target_accept=HMC["target_accept"], adapt=HMC["adapt"], n_chains=N_CHAINS, chain_offset=rank * N_CHAINS,
                         seed=seed0 + 7, keep_on_device=True, want_u=True)
    st = res["stats"]
    hmc_ms = st["kernel_ms"]
    # columns are chain-major (column = chain * per + draw): [q, draw, chain] in Fortran order -> [q, chain, draw]
    Uall = res["u"].reshape(Q, per, N_CHAINS, order="F").transpose(0, 2, 1)[:, :, 1:]   # drop each chain's column 0 (warm-up end state)
    ess = np.array([ess_chains(Uall[q]) for q in range(Q)])
    n_post = N_CHAINS * (per - 1)
    extra["hmc"] = {"u_samples_per_s": sum_over_ranks(N_CHAINS * per / (hmc_ms * 1e-3)),
                    "ess_per_s_min": sum_over_ranks(float(ess.min()) / (hmc_ms * 1e-3)),
                    "ess_per_s_median": sum_over_ranks(float(np.median(ess)) / (hmc_ms * 1e-3)),
                    "ess_over_n_median": float(np.median(ess) / n_post), "accept_rate": st["accept_rate"],
                    "step_size_mean": st["step_size_mean"], "steps_mean": st["steps_mean"], "ms": hmc_ms,
                    "leapfrog_per_s": sum_over_ranks(st["leapfrog_total"] / (hmc_ms * 1e-3))}
    hmc_flops = st["leapfrog_total"] * 4.0 * cfg["n"] * Q                     # 4 n Q per leapfrog step per chain
    hmc_tflops = hmc_flops / (hmc_ms * 1e-3) / 1e12
    if st["kernel_variant"] == 3:
        exec_per_step = 4.0 * st["zl_nonzeros"] + SPARSE_ROW_FLOPS[cfg["family"]] * st["rows_used"] + 8.0 * Q
    else:
        exec_per_step = 4.0 * st["rows_used"] * Q                             # dense contraction on the rows the kernel ran on
    hmc_exec_tflops = st["leapfrog_total"] * exec_per_step / (hmc_ms * 1e-3) / 1e12
    mdl.use_device_u()
    # E-step evaluations/s on the step's own zd (40 MB, L2-resident between evaluations) and cold (L2 flushed before each)
    ctx.timer_start(); mdl.log_likelihood_batch(Bst, np.ones(N_HESS)); t_b = ctx.timer_stop()
    cold = []
    for k in range(8):
        ctx.flush_l2(); ctx.sync()
        ctx.timer_start(); mdl.log_likelihood(Bst[:, k], 1.0); cold.append(ctx.timer_stop())
    extra["estep"] = {"loglik_evals_per_s_batched": N_HESS / (t_b * 1e-3), "loglik_evals_per_s_single_cold": 1e3 / float(np.median(cold)),
                      "m": M_PER_GPU * world, "note": "m=10^4 x n=500: 40 MB, below launch latency; see roofline_estep for the streaming rate"}
    ctx.timer_start(); [cv.loglik_model(Tst[:, k], mdl) for k in range(64)]; t_d = ctx.timer_stop()
    extra["mvn_ll_evals_per_s"] = 64 / (t_d * 1e-3)
    ctx.timer_start(); cv.loglik_model_batch(Tst, mdl); t_db = ctx.timer_stop()
    extra["mvn_ll_evals_per_s_batched"] = N_HESS / (t_db * 1e-3)
    ctx.timer_start(); [mdl.mcnr(beta, 1.0) for _ in range(16)]; t_n = ctx.timer_stop()
    extra["mcnr_steps_per_s"] = 16 / (t_n * 1e-3)

    # ---- BASELINE.json's large configurations at their stated size, with in-run oracle parity (tools/bench_configs.py) ----
    configs = None
    if not args.no_configs:
        import oracle as orc
        orc.build()
        if world > 1:
            orc.set_threads(max(1, (os.cpu_count() or world) // world))      # torchrun exports OMP_NUM_THREADS=1; the checker may use this rank's share
        configs = {"scaling": "strong: m (sample columns) and chains are split over the ranks; C3 (m = 250) does not shard and runs at N = 1 only",
                   "size": args.configs_size}
        for nm in (["C3"] if world == 1 else []) + ["C4", "C5"]:
            t0c = time.perf_counter()
            configs[nm] = bc.run_config(nm, env, size=args.configs_size, oracle=orc)
            configs[nm]["wall_s"] = round(time.perf_counter() - t0c, 1)
            if rank == 0:
                print("[bench] %s done in %.1f s, parity ok = %s" % (nm, configs[nm]["wall_s"], configs[nm]["parity"]["ok"]), file=sys.stderr, flush=True)

    # every collective of this run is behind us: tear the process group down on all ranks together, before rank 0's solo work
    barrier()
    if dist is not None:
        dist.destroy_process_group()
        dist = None
    peaks, peak_src = load_peaks()
    roofline_estep = None
    roofline_sat = None
    roofline_many = None
    if rank == 0:
        # HBM roofline probe of the E-step kernels: same model, 1 GB of zd (> 4 x L2), 8 evaluations in one batch
        mbig = 250_000
        rngb = np.random.default_rng(5)
        Ubig = np.asfortranarray(cfg["L"] @ rngb.standard_normal((Q, mbig)))
        # rank 0 only: on a context WITHOUT communicator (a collective issued by one rank alone would never return)
        ctx1 = g.Context(local_rank) if world > 1 else ctx
        mdl2 = g.Model(ctx1, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
        # the streaming kernels a design WITHOUT repeated rows runs (C3-C5): one row of zd per observation.  C2 itself repeats every row of
        # [X | Z] ten times and its E-step runs on the 50 distinct rows by default (`aggregated` below: the same evaluations on a tenth of the bytes)
        g.estep_set_row_aggregation(False)
        try:
            mdl2.set_u(Ubig)
        finally:
            g.estep_set_row_aggregation(True)
        g.estep_set_multi(False)                     # one launch per evaluation: every launch streams the whole 1 GB
        try:
            mdl2.log_likelihood_batch(Bst[:, :4], np.ones(4))
            ctx1.timer_start(); mdl2.log_likelihood_batch(Bst[:, :8], np.ones(8)); t_big = ctx1.timer_stop() / 8
        finally:
            g.estep_set_multi(True)
        mdl2.log_likelihood_batch(Bst[:, :64], np.ones(64))    # the batched kernel: 8 evaluations per pass over the matrix
        ctx1.timer_start(); mdl2.log_likelihood_batch(Bst[:, :64], np.ones(64)); t_multi = ctx1.timer_stop() / 64
        bytes_ll = 8.0 * cfg["n"] * mbig + 16.0 * cfg["n"]
        mdl2.mcnr(beta, 1.0)
        ctx1.timer_start(); [mdl2.mcnr(beta, 1.0) for _ in range(4)]; t_nr = ctx1.timer_stop() / 4
        bytes_nr = 8.0 * cfg["n"] * mbig + 8.0 * cfg["n"] * (P + 2)
        roofline_estep = {"kernel": "loglik_logit_factor_kernel (binomial-logit log-likelihood on the factor matrix, one evaluation per launch)", "bound": "hbm", "achieved": bytes_ll / (t_big * 1e-3) / 1e9,
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": bytes_ll / (t_big * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "bytes_per_launch": bytes_ll, "traffic": LOGLIK_DRAM_BYTES_PER_LAUNCH, "ms": t_big, "zd_bytes": 8.0 * cfg["n"] * mbig, "peak_source": peak_src,
                          "batched": {"ms_per_eval": t_multi, "evals_per_s": 1e3 / t_multi,
                                      "note": "loglik_logit_factor_multi_kernel, 64 evaluations in one launch: the matrix is read once per 8 evaluations "
                                              "(algorithmic bytes / 8 per evaluation), bound by FP64 work"},
                          "mcnr": {"achieved": bytes_nr / (t_nr * 1e-3) / 1e9, "frac": bytes_nr / (t_nr * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms": t_nr,
                                   "traffic": MCNR_DRAM_BYTES_PER_LAUNCH,
                                   "note": "whole mcnr() call (x'beta, TMA pass, tail, 7 doubles back to the host); the TMA pass alone: 171 us = 5.83 TB/s under ncu"}}
        mdl2.close()
        # the same evaluations with the E-step on the 50 distinct rows (the default for this model): zd is 100 MB instead of 1 GB
        mdl2 = g.Model(ctx1, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
        mdl2.set_u(Ubig)
        rows_agg = mdl2.estep_rows()
        mdl2.log_likelihood_batch(Bst[:, :64], np.ones(64))
        ctx1.timer_start(); mdl2.log_likelihood_batch(Bst[:, :64], np.ones(64)); t_agg = ctx1.timer_stop() / 64
        mdl2.mcnr(beta, 1.0)
        ctx1.timer_start(); [mdl2.mcnr(beta, 1.0) for _ in range(4)]; t_nr_agg = ctx1.timer_stop() / 4
        roofline_estep["aggregated"] = {"rows": rows_agg, "zd_bytes": 8.0 * rows_agg * mbig, "loglik_ms_per_eval_batched": t_agg, "mcnr_ms": t_nr_agg,
                                        "speedup_loglik_vs_batched_dense": t_multi / t_agg, "speedup_mcnr": t_nr / t_nr_agg,
                                        "note": "SURVEY 8f N2: observations sharing their row of [X | Z] share eta; zd and every E-step kernel run on the distinct rows"}
        mdl2.close()
        # the timed step's sampler kernel with every scheduler holding several warps (4000 chains): its throughput regime
        mdl4 = g.Model(ctx1, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
        for rep in range(2):
            r4 = mdl4.hmc_sample(L, beta, 1.0, warmup=100, nsamp_per_chain=4, lam=HMC["lam"], max_steps=HMC["max_steps"],
                                 target_accept=HMC["target_accept"], adapt=HMC["adapt"], n_chains=4000, seed=seed0 + 98, keep_on_device=True,
                                 want_u=False)
        s4 = r4["stats"]
        roofline_many = {"chains": 4000, "ms": s4["kernel_ms"], "leapfrog_per_s": s4["leapfrog_total"] / (s4["kernel_ms"] * 1e-3),
                         "kernel_variant": s4["kernel_variant"]}
        mdl4.close()
        # the sampler kernel as a dense tensor kernel: row aggregation off, every SM busy on 8 tiles per warp (1184 chains = 148 groups,
        # one CTA each) — what the DMMA path reaches on a model without repeated rows
        g.hmc_set_row_aggregation(False)
        g.hmc_set_variant(2)
        mdl3 = g.Model(ctx1, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
        for rep in range(2):
            r3 = mdl3.hmc_sample(L, beta, 1.0, warmup=100, nsamp_per_chain=4, lam=HMC["lam"], max_steps=HMC["max_steps"],
                                 target_accept=HMC["target_accept"], adapt=HMC["adapt"], n_chains=1184, seed=seed0 + 99, keep_on_device=True,
                                 want_u=False)
        s3 = r3["stats"]
        sat_tflops = s3["leapfrog_total"] * 4.0 * cfg["n"] * Q / (s3["kernel_ms"] * 1e-3) / 1e12
        roofline_sat = {"chains": 1184, "ms": s3["kernel_ms"], "achieved": sat_tflops, "frac": sat_tflops / FP64_DMMA_PEAK_TFLOPS,
                        "kernel": "hmc_fused_kernel<binomial-logit, KS=13> (dense on-chip sampler: eta = ZL v and grad = ZL^T r(eta) as FP64 DMMA)",
                        "note": "the DENSE on-chip kernel (forced; the dispatcher picks the structure-aware one for this model) without row aggregation, "
                                "1184 chains (148 CTAs x 8 chains, no cluster split), 104 proposals: the DMMA path on all 500 rows, 4 n Q flop per leapfrog "
                                "step and chain against the measured DMMA peak — what a model with a dense Z L gets; not the timed step"}
        mdl3.close()
        g.hmc_set_variant(0)
        g.hmc_set_row_aggregation(True)

    if rank != 0:
        return
    cpu = None
    comparators = None
    if world == 1 and not args.no_cpu_baseline:
        ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        t_cpu, det, kind, sample = cpu_arm(cfg, ncores, args.quick)
        if not args.quick:
            det.update(cpu_hoisted_and_full_m(cfg, det, ncores))
        cpu = {"value": det["m_sample"] / t_cpu, "unit": "u-samples/s", "cores": det["threads"], "kind": kind, "sample": sample, "sample_s": t_cpu,
               "hoisted_value": det.get("hoisted_value"), "detail": det}
        comparators = vendor_comparators()
    line = {"metric": "u-samples/s through sampler + E-step", "value": value, "unit": "u-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload,
            "e2e": {"value": e2e_val, "unit": "u-samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "calls": "gmb_mcmc_sample + gmb_mcml_optim(mcnr) + gmb_mcml_hess with host buffers", "s_per_step": e2e_s,
                    "parts_s": {k: float(np.mean(v)) for k, v in e2e_parts.items()}},
            "gpu_launches": int(launches), "clocks": clk, "wall_s": wall,
            "roofline": sampler_roofline(st, hmc_exec_tflops, hmc_tflops, exec_per_step, cfg, Q, roofline_sat, roofline_many),
            "roofline_estep": roofline_estep, "cpu_baseline": cpu, "step_parity": step_parity, "configs": configs,
            "vendor_fp64_comparators": comparators}
    line.update(extra)
    emit(line)


if __name__ == "__main__":
    main()
