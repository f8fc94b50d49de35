#!/usr/bin/env python
"""Short program for `ncu --set full` on the sampler kernels (C2 model, 62 proposals each):
  launch 1: the structure-aware kernel as the bench runs it (1000 chains, 50 distinct rows, 150 non-zeros of Z L; one warp per chain)
  launch 2: the dense on-chip kernel, forced, 1184 chains without row aggregation (the DMMA path on all 500 rows, 148 CTAs)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth
ctx = g.Context(0); cfg = synth.config2(m=64)
for label, agg, variant, nch in (("structure-aware", True, 0, 1000), ("dense on-chip", False, 2, 1184)):
    g.hmc_set_row_aggregation(agg); g.hmc_set_variant(variant)
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
    out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=60, nsamp_per_chain=2, lam=5.0, max_steps=100, target_accept=0.95,
                         n_chains=nch, seed=3, keep_on_device=True, want_u=False)
    print(label, out["stats"])
    mdl.close()
g.hmc_set_row_aggregation(True); g.hmc_set_variant(0)
