#!/usr/bin/env python
"""Short program for `ncu --set full`: launches each hot kernel a few times at a size that shows its steady state.
  - loglik_kernel / mcnr_pass1_kernel on 1 GB of zd (C2 model, m = 250000)  -> HBM-bound E-step stream
  - hmc_fused_kernel on the C2 model, 1184 chains, 30 proposals              -> on-chip sampler
  - dgemm_kernel (zd = Z u, two-GEMM sampler variant, large-block Cholesky)  -> DMMA contractions
Run it plainly first (it must exit 0), then under ncu (see profiles/README.md)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
g.hmc_set_variant(2)        # this tool studies the dense on-chip kernel (the dispatcher alone picks the structure-aware one for C2)
from glmmrmcml_b200 import synth

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = g.Context(0)
cfg = synth.config2(m=64)
rng = np.random.default_rng(0)
if which in ("all", "estep"):
    m = 250_000
    U = np.asfortranarray(cfg["L"] @ rng.standard_normal((cfg["Q"], m)))
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
    mdl.set_u(U)
    for _ in range(3):
        mdl.log_likelihood(cfg["beta"], 1.0)
    for _ in range(2):
        mdl.mcnr(cfg["beta"], 1.0)
    B = np.asfortranarray(np.repeat(cfg["beta"][:, None], 16, axis=1) * (1 + 1e-6 * np.arange(16))[None, :])
    mdl.log_likelihood_batch(B, np.ones(16))        # the batched kernel: 16 evaluations in one launch, 8 per pass over the matrix
    mdl.close()
if which in ("all", "hmc"):
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
    out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=25, nsamp_per_chain=5, lam=5.0, max_steps=100, target_accept=0.95,
                         n_chains=1184, seed=3, keep_on_device=True, want_u=False)
    print("fused", out["stats"])
    mdl.close()
if which in ("all", "cov"):
    c3 = synth.config3(nloc=4000, m=256)
    cv = g.Covariance(ctx, c3["cov"], c3["data"], c3["eff_range"])
    print("mvn_ll C3 n=4000", cv.loglik(c3["theta"], c3["U"]))
    cv.close()
ctx.close()
print("done")
