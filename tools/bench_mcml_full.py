#!/usr/bin/env python
"""Whole fits through the reference-named entry point mcml_full (src/mcml_full.cpp:41-148) on BASELINE.json's configs[0] and configs[1]:
C1 (README cluster RCT, MCEM, m = 250, tol = 5e-3) and C2 (gr(cl)*ar1(t), MCNR, m = 10^4), host buffers in, estimates out.
One JSON line per fit:  python tools/bench_mcml_full.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth

ctx = g.Context(0); ctx.make_default()
for name, cfg, mcnr, m, tol, chains in (("C1 MCEM m=250 tol=5e-3", synth.config1(m=4), False, 250, 5e-3, 0),
                                        ("C2 MCNR m=10^4 tol=1e-2", synth.config2(m=4), True, 10_000, 1e-2, 1000)):
    start = np.concatenate([cfg["beta"] * 0.0 + 0.1, cfg["theta"] * 0.0 + 0.5, [1.0]])      # away from the truth
    for rep in range(2):
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        fit = g.mcml_full(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start,
                          mcnr=mcnr, m=m, maxiter=30, warmup=500, tol=tol, verbose=False, lam=5.0, maxsteps=100, target_accept=0.95,
                          n_chains=chains, seed=7 + rep)
        dt = time.perf_counter() - t0
    print(json.dumps({"fit": name, "seconds": round(dt, 4), "iterations": fit["iter"], "converged": fit["converged"], "launches": ctx.launch_count - l0,
                      "beta": np.round(fit["beta"], 3).tolist(), "theta": np.round(fit["theta"], 3).tolist(),
                      "true_beta": np.round(cfg["beta"], 3).tolist(), "true_theta": np.round(cfg["theta"], 3).tolist()}), flush=True)
