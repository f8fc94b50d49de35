#!/usr/bin/env python
"""Short program for `ncu --set full` on the on-chip sampler (C2 model, 62 proposals each):
  launch 1: 1000 chains with row aggregation (the bench's configuration: 50 distinct rows, one CTA per group of 8 chains)
  launch 2: 1184 chains without row aggregation (the kernel as a dense tensor kernel: 500 rows, 148 CTAs)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth
ctx = g.Context(0); cfg = synth.config2(m=64)
for agg, nch in ((True, 1000), (False, 1184)):
    g.hmc_set_row_aggregation(agg)
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
    out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=60, nsamp_per_chain=2, lam=5.0, max_steps=100, target_accept=0.95,
                         n_chains=nch, seed=3, keep_on_device=True, want_u=False)
    print("fused", "aggregated" if agg else "dense", out["stats"])
    mdl.close()
g.hmc_set_row_aggregation(True)
