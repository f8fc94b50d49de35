"""Profiling drivers for the round-2 kernels (run under ncu by tools/gpu_profile_r02.sh):
  hmc     the bench's sampler draw on C2 (lane-per-component kernel), shortened to 62 proposals
  estep   log-likelihood and MCNR step on 1 GB of zd / F (binomial, C2 model, m = 250000), fp64 and fp32 storage
  gemm    zd = Z u as a dense 8192 x 4096 x 16384 contraction: fp64 DMMA (TMA kernel) and fp32 mode (tcgen05 3xTF32)
  chol    one factorisation + mvn_ll of a dense 5000 x 5000 exponential-covariance block, 10^4 sample columns"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth

what = sys.argv[1]
ctx = g.Context(0)
rng = np.random.default_rng(1)
if what == "hmc":
    cfg = synth.config2(m=8)
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
    for rep in range(2):
        r = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=52, nsamp_per_chain=9, lam=5.0, max_steps=100, target_accept=0.95, n_chains=1000, seed=3 + rep,
                           keep_on_device=True, want_u=False)
    print(r["stats"])
elif what == "estep":
    cfg = synth.config2(m=8)
    U = np.asfortranarray(cfg["L"] @ rng.standard_normal((cfg["Q"], 250_000)))
    g.estep_set_row_aggregation(False)           # the per-observation streaming kernels (what C3-C5 run): 1 GB per pass
    for prec in ("fp64", "fp32"):
        mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit", precision=prec)
        mdl.set_u(U)
        for rep in range(2):
            mdl.log_likelihood(cfg["beta"] * (1 + 0.01 * rep), 1.0); mdl.mcnr(cfg["beta"], 1.0)
        mdl.close()
    g.estep_set_row_aggregation(True)            # C2's default: the same evaluations on its 50 distinct rows (100 MB), 64 evaluations per batch
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
    mdl.set_u(U)
    B = np.asfortranarray(cfg["beta"][:, None] * (1 + 1e-3 * np.arange(64))[None, :])
    for rep in range(2):
        mdl.log_likelihood_batch(B, np.ones(64)); mdl.mcnr(cfg["beta"], 1.0)
    mdl.close()
elif what == "gemm":
    n, Q, m = 8192, 4096, 16384
    X = np.asfortranarray(np.ones((n, 1))); Z = np.asfortranarray(rng.standard_normal((n, Q)) / np.sqrt(Q))
    U = np.asfortranarray(rng.standard_normal((Q, m))); y = rng.standard_normal(n)
    for prec in ("fp64", "fp32"):
        mdl = g.Model(ctx, X, Z, y, "gaussian", "identity", precision=prec)
        mdl.set_u(U); mdl.rebuild_zd()
        mdl.close()
elif what == "chol":
    nloc = 5000
    xy = rng.random((nloc, 2))
    cov = np.array([[0, nloc, 13, 2, 0]], dtype=np.int32); data = np.concatenate([xy[:, 0], xy[:, 1]])
    cv = g.Covariance(ctx, cov, data, np.zeros(1))
    U = np.asfortranarray(rng.standard_normal((nloc, 10_000)))
    th = np.array([0.25, 0.1])
    cv.loglik(th, U)
    cv.loglik(th * 1.001, U)
print("done", what)
