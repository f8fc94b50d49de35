#!/usr/bin/env python
"""Generates tests/golden/*.npz from oracle/_ref/libref.so — the REFERENCE'S OWN headers (inst/include/glmmrmcml/*.h)
compiled unmodified against oracle/shim (see oracle/ref_driver.cpp).  Run in the dev container, where /root/reference
exists:   make -C oracle ref && python tests/golden/make_golden.py

Each fixture stores the inputs (so the tests do not depend on the synthetic-data generator staying unchanged) and the
reference outputs: E-step log-likelihood at several (beta, sigma), the MCNR step, mvn_ll at several theta, log_prob /
log_grad at several whitened states, a short HMC run driven by the shared Philox stream, and the three objective
functors of likelihood.h at one point.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from glmmrmcml_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

CASES = {
    "C1_binomial_gr": lambda: synth.config1(m=24),
    "C2_binomial_gr_ar1": lambda: synth.config2(m=24),
    "C3_gaussian_fexp": lambda: synth.config3(nloc=48, m=16),
    "C4_poisson_gr_ar1": lambda: synth.config4(ncl=12, nt=6, k=2, m=20),
    "C5_binomial_fexp": lambda: synth.config5(nloc=40, nobs=3, m=16),
}
HMC = dict(warmup=14, nsamp=9, lam=0.05, maxsteps=12, target_accept=0.9, seed=20221208, chain=3)


def main():
    assert ref.available(), "build oracle/_ref first: make -C oracle ref"
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, make in CASES.items():
        cfg = make()
        fam, link = cfg["family"], cfg["link"]
        X, Z, y, U, L = cfg["X"], cfg["Z"], cfg["y"], cfg["U"], cfg["L"]
        cov, data, eff, theta, beta = cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], cfg["beta"]
        rng = np.random.default_rng(99)
        betas = beta[:, None] + 0.2 * rng.standard_normal((beta.size, 4)); betas[:, 0] = beta
        sigmas = np.array([1.0, 0.7, 1.3, 2.0])
        ll = np.array([ref.loglik(X, Z, U, y, betas[:, k], sigmas[k], fam, link) for k in range(4)])
        thetas = theta[:, None] * np.array([[1.0, 1.2, 0.8], [1.0, 0.9, 1.1]])[: theta.size]
        mvn = np.array([ref.mvn_loglik(cov, data, eff, thetas[:, k], U) for k in range(3)])
        logdet = np.array([ref.logdet(cov, data, eff, thetas[:, k]) for k in range(3)])
        start = np.concatenate([beta, theta, [1.0]])
        nr_beta, nr_sigma = ref.mcnr(cov, data, eff, X, Z, U, y, fam, link, start)
        V = 0.6 * rng.standard_normal((cfg["Q"], 3))
        lp = np.array([ref.log_prob(X, Z, L, y, beta, 0.9, fam, link, V[:, k]) for k in range(3)])
        lg = np.stack([ref.log_grad(X, Z, L, y, beta, 0.9, fam, link, V[:, k]) for k in range(3)], axis=1)
        hm_u, hm_st = ref.mcmc_sample(X, Z, L, y, beta, fam, link, HMC["warmup"], HMC["nsamp"], HMC["lam"], 1.0, HMC["maxsteps"],
                                      HMC["target_accept"], HMC["seed"], HMC["chain"])
        obj = ref.objectives(cov, data, eff, X, Z, U, y, fam, link, np.concatenate([beta, theta]), 1.0)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), family=fam, link=link, X=X, Z=Z, y=y, U=U, L=L, cov=cov, data=data,
                            eff_range=eff, theta=theta, beta=beta, betas=betas, sigmas=sigmas, loglik=ll, thetas=thetas, mvn_ll=mvn,
                            logdet=logdet, mcnr_beta=nr_beta, mcnr_sigma=nr_sigma, V=V, log_prob=lp, log_grad=lg, hmc_u=hm_u,
                            hmc_accept=hm_st["accept"], hmc_eps=hm_st["eps"], hmc_steps=hm_st["steps"], objectives=obj,
                            hmc_settings=np.array([HMC["warmup"], HMC["nsamp"], HMC["lam"], HMC["maxsteps"], HMC["target_accept"], HMC["seed"], HMC["chain"]]))
        print(name, "loglik", ll[0], "mvn", mvn[0], "hmc accept", hm_st["accept"], "eps", hm_st["eps"])
    # scalar family terms on a grid (moremaths.h:26-102) for the three in-scope families
    ys = {1: [0, 1, 2, 5, 17], 3: [0, 1], 7: [-1.3, 0.0, 2.5]}
    rows = []
    for fl, yl in ys.items():
        for yv in yl:
            for eta in (-6.0, -1.5, -0.2, 0.0, 0.3, 2.0, 7.5):
                for sg in (0.5, 1.0, 2.5):
                    rows.append([fl, yv, eta, sg, ref.family_ll(yv, eta, sg, fl)])
    np.savez_compressed(os.path.join(out_dir, "family_terms.npz"), table=np.array(rows),
                        log_factorial=np.array([[k, ref.log_factorial_approx(k)] for k in range(0, 60)]))
    print("family table rows", len(rows))


if __name__ == "__main__":
    main()
