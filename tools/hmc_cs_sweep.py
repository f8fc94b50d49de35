#!/usr/bin/env python
"""On-chip sampler timing at cluster sizes 1/2/4 on the C2 model (500 and 1184 chains)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
g.hmc_set_variant(2)        # this tool studies the dense on-chip kernel (the dispatcher alone picks the structure-aware one for C2)
from glmmrmcml_b200 import synth
ctx = g.Context(0); cfg = synth.config2(m=64)
mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
for nch in (500, 1184):
    for cs in (1, 2, 4, 0):
        g.hmc_set_cluster_size(cs)
        for rep in range(2):
            out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=100, nsamp_per_chain=20, lam=5.0, max_steps=100, target_accept=0.95,
                                 n_chains=nch, seed=3, keep_on_device=True, want_u=False)
        st = out["stats"]
        print(f"chains {nch} cs {cs}: kernel_ms {st['kernel_ms']:.2f}  us/leapfrog-step {st['kernel_ms']*1e3/(120*st['steps_mean']):.3f}  accept {st['accept_rate']:.4f} eps {st['step_size_mean']:.5f}")
