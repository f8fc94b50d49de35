#!/usr/bin/env python
"""Short program for `ncu --set full` on the on-chip sampler: C2 model, 500 chains (cluster size chosen automatically), 12 proposals."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 500
ctx = g.Context(0); cfg = synth.config2(m=64)
mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=60, nsamp_per_chain=2, lam=5.0, max_steps=100, target_accept=0.95,
                     n_chains=nch, seed=3, keep_on_device=True, want_u=False)
print("fused", out["stats"])
