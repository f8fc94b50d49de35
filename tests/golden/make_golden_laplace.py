#!/usr/bin/env python
"""Generates tests/golden/LA_*.npz: outputs of the reference's own Laplace functors and Newton step (likelihood.h:112-230,
mcmloptim.h:238-293) from oracle/_ref/libref.so — the reference headers compiled where they lie under /root/reference against
oracle/shim — on small seeded inputs.  Run in the dev container (needs /root/reference): python tests/golden/make_golden_laplace.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from oracle import ref
from glmmrmcml_b200 import synth

oracle.build()
CASES = {
    "LA_binomial_gr_ar1": (synth.config2(m=4, seed=11, ncl=8, nt=4, nind=6), "binomial", "logit"),
    "LA_poisson_gr_ar1": (synth.config4(ncl=12, nt=5, k=3, m=4), "poisson", "log"),
    "LA_gaussian_fexp": (synth.config3(nloc=40, m=4), "gaussian", "identity"),
}
for name, (cfg, fam, link) in CASES.items():
    rng = np.random.default_rng(3)
    beta = cfg["beta"] + 0.05 * rng.standard_normal(cfg["P"])
    theta = cfg["theta"] * np.array([1.1, 0.9])[: cfg["theta"].size]
    v = 0.4 * rng.standard_normal(cfg["Q"])
    sigma = 0.8 if fam == "gaussian" else 1.0
    out = {}
    for use_l in (0, 1):
        r = ref.la_objectives(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["X"], cfg["Z"], cfg["y"], fam, link, beta, theta, v, sigma, bool(use_l))
        out[f"obj_{use_l}"] = np.array([r["la"], r["la_cov"], r["la_btheta"]])
        out[f"beta_nr_{use_l}"] = r["beta_nr"]; out[f"v_nr_{use_l}"] = r["v_nr"]; out[f"sigma_nr_{use_l}"] = np.array(r["sigma_nr"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), family=fam, link=link, X=cfg["X"], Z=cfg["Z"], y=cfg["y"],
                        cov=cfg["cov"], data=cfg["data"], eff_range=cfg["eff_range"], beta=beta, theta=theta, v=v, sigma=np.array(sigma), **out)
    print(name, out["obj_0"])
