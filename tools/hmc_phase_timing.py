#!/usr/bin/env python
"""Per-phase cycle breakdown of the on-chip sampler's leapfrog step.  Needs the instrumented library:
    make -C glmmrmcml_b200/csrc BUILD=build_timing LIB=../libglmmrmcml_b200_timing.so EXTRA=-DGMB_FUSED_TIMING
    GMB_LIB=glmmrmcml_b200/libglmmrmcml_b200_timing.so python tools/hmc_phase_timing.py [chains] [cluster size]
The counters of the last launch are printed to stderr at exit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
g.hmc_set_variant(2)        # this tool studies the dense on-chip kernel (the dispatcher alone picks the structure-aware one for C2)
from glmmrmcml_b200 import synth
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 500
cs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = g.Context(0); cfg = synth.config2(m=64)
mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
g.hmc_set_cluster_size(cs)
for rep in range(2):
    out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=100, nsamp_per_chain=20, lam=5.0, max_steps=100, target_accept=0.95,
                         n_chains=nch, seed=3, keep_on_device=True, want_u=False)
st = out["stats"]
print(f"chains {nch} cs {cs}: kernel_ms {st['kernel_ms']:.2f} us/step {st['kernel_ms']*1e3/(120*st['steps_mean']):.3f} steps/chain {120*st['steps_mean']:.0f}")
