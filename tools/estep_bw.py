#!/usr/bin/env python
"""HBM streaming rate of the E-step kernels (log-likelihood, MCNR sums) on 1 GB of zd, per family."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth
ctx = g.Context(0)
fams = sys.argv[1:] or ["binomial", "poisson", "gaussian"]
for fam in fams:
    if fam == "binomial":
        cfg = synth.config2(m=64); link = "logit"; mbig = 250_000
    elif fam == "poisson":
        cfg = synth.config4(ncl=50, nt=10, k=1, m=64); link = "log"; mbig = 250_000
    else:
        cfg = synth.config3(nloc=500, m=64); link = "identity"; mbig = 250_000
    n, Q, P = cfg["n"], cfg["Q"], cfg["P"]
    rng = np.random.default_rng(5)
    U = np.asfortranarray(cfg["L"] @ rng.standard_normal((Q, mbig)))
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], fam, link)
    mdl.set_u(U)
    B = np.asfortranarray(np.repeat(cfg["beta"][:, None], 8, axis=1) + 1e-6 * np.arange(8)[None, :])
    g.estep_set_rowstats(False)                      # the stream itself (poisson / gaussian default to O(n) row-statistic evaluations)
    g.estep_set_multi(False)                         # ... one launch per evaluation (binomial batches default to 8 evaluations per pass)
    mdl.log_likelihood_batch(B[:, :4], np.ones(4))
    ctx.timer_start(); mdl.log_likelihood_batch(B, np.ones(8)); t = ctx.timer_stop() / 8
    g.estep_set_rowstats(True); g.estep_set_multi(True)
    mdl.log_likelihood_batch(B[:, :4], np.ones(4))
    ctx.timer_start(); mdl.log_likelihood_batch(B, np.ones(8)); t_default = ctx.timer_stop() / 8
    by = 8.0 * n * mbig + 16.0 * n
    mdl.mcnr(cfg["beta"], 1.0)
    ctx.timer_start(); [mdl.mcnr(cfg["beta"], 1.0) for _ in range(4)]; tn = ctx.timer_stop() / 4
    byn = 8.0 * n * mbig + 8.0 * n * (P + 2)
    print(f"{fam:9s} n={n} m={mbig}: loglik stream {t:.3f} ms {by/t/1e6:7.0f} GB/s ({by/t/1e6/6539.9*100:.1f}% of 6539.9), default path {t_default*1e3:.1f} us/eval | mcnr {tn:.3f} ms {byn/tn/1e6:7.0f} GB/s ({byn/tn/1e6/6539.9*100:.1f}%)")
    mdl.close()
