#!/usr/bin/env python
"""The reference's OWN headers (oracle/_ref/libref_omp.so: inst/include/glmmrmcml/*.h compiled where they lie against oracle/shim, OpenMP
pragmas on) timed piece by piece beside the oracle's FAITHFUL port on the SAME small sample of the C2 step.  bench.py runs this in a
subprocess (the OpenMP build of mcmloptim::mcnr races, SURVEY §5 — its numbers are timings, never compared) and reports the ratios in
`cpu_baseline.detail.reference_headers`: they show that the port, whose time is the CPU arm's value, is not slower than the headers it restates
(the stand-in Eigen of oracle/shim evaluates products with plain loops).  Also: the port's pieces on ONE thread (`port_single_thread`,
`port_parallel_speedup` — a single HMC chain of this size does not speed up with threads).  Test infrastructure; prints one JSON line.
Usage: ref_headers_timing.py [threads] [m_sample] [proposals]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import oracle
    from oracle import ref
    from glmmrmcml_b200 import synth
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else len(os.sched_getaffinity(0))
    m_s = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    props = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    if not ref.timing_available():
        print(json.dumps({"unavailable": "oracle/_ref/libref_omp.so not built (needs the reference tree at build time)"}))
        return
    oracle.build()
    oracle.set_threads(threads)            # one libgomp per process: this also sets the team size of libref_omp.so
    ref.use_timing_build()
    cfg = synth.config2(m=m_s)
    X, Z, y, beta, theta, L, U = cfg["X"], cfg["Z"], cfg["y"], cfg["beta"], cfg["theta"], cfg["L"], cfg["U"]
    fam, link = cfg["family"], cfg["link"]
    fl = oracle.flink(fam, link)
    cov = (cfg["cov"], cfg["data"], cfg["eff_range"])
    ZL = Z @ L
    xb = X @ beta
    warm = props // 4
    out = {"threads": threads, "m_sample": m_s, "hmc_proposals": props}

    def clock(f, reps=1):
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        return (time.perf_counter() - t0) / reps

    # sampler: mcmcRunHMC::sample (mhmcmc.h:121-157) on the R-default trajectory settings of the bench
    t_ref = clock(lambda: ref.mcmc_sample(X, Z, L, y, beta, fam, link, warm, props - warm, 5.0, 1.0, 100, 0.95, 12345))
    t_port = clock(lambda: oracle.hmc_chain(ZL, L, xb, y, 1.0, fl, warm, props - warm, 5.0, 100, 0.95, 12345, want_u=False))
    out["hmc_s_per_proposal"] = {"headers": t_ref / props, "port": t_port / props}
    # MCNR step (mcmloptim.h:198-236 through mcml_optim's construction, src/mcml_optim.cpp:48-62)
    start = np.concatenate([beta, theta, [1.0]])
    t_ref = clock(lambda: ref.mcnr(*cov, X, Z, U, y, fam, link, start))
    t_port = clock(lambda: oracle.mcnr(X, Z, U, y, beta, 1.0, fl, faithful=True))
    out["mcnr_s"] = {"headers": t_ref, "port": t_port}
    # log_likelihood() on an existing model (mcmlmodel.h:284-304, Z u product included): difference of 1 and 1 + r evaluations
    r = 8
    t1 = clock(lambda: ref.loglik_reps(X, Z, U, y, beta, 1.0, fam, link, 1))
    t2 = clock(lambda: ref.loglik_reps(X, Z, U, y, beta, 1.0, fam, link, 1 + r))
    t_port = clock(lambda: oracle.loglik_faithful(X, Z, U, y, beta, 1.0, fl), reps=r)
    out["loglik_s_per_eval"] = {"headers": max(t2 - t1, 1e-9) / r, "port": t_port}
    # MCMLDmatrix::loglik (mcmldmatrix.h:23-41)
    t_ref = clock(lambda: ref.mvn_loglik(*cov, theta, U), reps=r)
    t_port = clock(lambda: oracle.mvn_loglik(*cov, theta, U, faithful=True), reps=r)
    out["mvn_ll_s_per_eval"] = {"headers": t_ref, "port": t_port}
    for k in ("hmc_s_per_proposal", "mcnr_s", "loglik_s_per_eval", "mvn_ll_s_per_eval"):
        out[k]["headers_over_port"] = out[k]["headers"] / out[k]["port"]
    # the reference's own ENTRY POINTS (src/mcml_full.cpp, src/mcml_optim.cpp compiled against the shim, OpenMP build): the three calls the GPU
    # arm's e2e figure makes — mcmc_sample, mcml_optim(mcnr = TRUE), mcml_hess — on the same small sample, R-default warm-up shortened like the sampler piece
    try:
        from oracle import refsrc
        if refsrc.timing_available():
            refsrc.use_timing_build()
            start = np.concatenate([beta, theta, [1.0]])
            ep = {"mcmc_sample_s_per_proposal": clock(lambda: refsrc.mcmc_sample(Z, L, X, y, beta, fam, link, warm, props - warm, 5.0, 1.0, 0, 500, 100, 0.95, seed=12345)) / props,
                  "mcml_optim_mcnr_s": clock(lambda: refsrc.mcml_optim(*cov, Z, X, y, U, fam, link, start, 0, True)),
                  "mcml_hess_s": clock(lambda: refsrc.mcml_hess(*cov, Z, X, y, U, fam, link, start[:-1], 1e-5, 0))}
            step = ep["mcmc_sample_s_per_proposal"] * (500 + m_s) + ep["mcml_optim_mcnr_s"] + ep["mcml_hess_s"]
            ep["step_s_at_m_sample"] = step
            ep["value"] = m_s / step
            ep["what"] = ("oracle/_ref/librefsrc_omp.so: the reference's own mcmc_sample + mcml_optim(mcnr) + mcml_hess bodies, one pass at m = m_sample "
                          "(sampler: measured per proposal x (500 warm-up + m_sample)); u-samples/s = m_sample / that time")
            out["entry_points"] = ep
    except Exception as e:
        out["entry_points"] = {"unavailable": str(e)[:200]}
    # the port once more on ONE thread (SURVEY §8d asks for both thread counts): how much of the CPU arm's speed is parallelism
    oracle.set_threads(1)
    p1 = max(40, props // 4)
    one = {"hmc_s_per_proposal": clock(lambda: oracle.hmc_chain(ZL, L, xb, y, 1.0, fl, p1 // 4, p1 - p1 // 4, 5.0, 100, 0.95, 12345, want_u=False)) / p1,
           "mcnr_s": clock(lambda: oracle.mcnr(X, Z, U, y, beta, 1.0, fl, faithful=True)),
           "loglik_s_per_eval": clock(lambda: oracle.loglik_faithful(X, Z, U, y, beta, 1.0, fl), reps=2),
           "mvn_ll_s_per_eval": clock(lambda: oracle.mvn_loglik(*cov, theta, U, faithful=True), reps=2)}
    out["port_single_thread"] = one
    out["port_parallel_speedup"] = {k: one[k] / out[k]["port"] for k in one}
    oracle.set_threads(threads)
    out["what"] = ("oracle/_ref/libref_omp.so = the reference's own headers compiled against oracle/shim (stand-in Eigen / Rcpp / glmmrBase), OpenMP on; "
                   "`port` = oracle FAITHFUL on the same inputs and thread count; ratio > 1: the headers are slower than the port the CPU arm times")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
