#!/usr/bin/env python
"""Times the three reference-named calls of bench.py's end-to-end step separately (host buffers in, host results out)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth

M, NCH = 10_000, 1000
cfg = synth.config2(m=M)
ctx = g.Context(0); ctx.make_default()
cv = g.Covariance(ctx, cfg["cov"], cfg["data"], cfg["eff_range"])
L = cv.genD(cfg["theta"], chol=True)
start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
fam, link = cfg["family"], cfg["link"]
for i in range(4):
    t0 = time.perf_counter()
    u = g.mcmc_sample(cfg["Z"], L, cfg["X"], cfg["y"], cfg["beta"], fam, link, 500, M - 1, 5.0, 1.0, 0, 500, 100, 0.95, n_chains=NCH, seed=100 + i)
    t1 = time.perf_counter()
    fit = g.mcml_optim(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], u, fam, link, start, 0, True)
    t2 = time.perf_counter()
    H = g.mcml_hess(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], u, fam, link,
                    np.concatenate([fit["beta"], fit["theta"]]), 1e-5, 0)
    t3 = time.perf_counter()
    print(f"rep {i}: mcmc_sample {1e3*(t1-t0):.1f} ms  mcml_optim(mcnr) {1e3*(t2-t1):.1f} ms  mcml_hess {1e3*(t3-t2):.1f} ms  launches so far {ctx.launch_count}")
print("theta", fit["theta"], "beta", fit["beta"])
