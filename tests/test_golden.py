"""Golden vectors produced by the reference's own headers (tests/golden/make_golden.py, oracle/_ref) checked against
(a) the CPU oracle — bit-exact or to rounding — and (b) the CUDA path through the C-ABI (gpu marker, 1e-10 relative)."""
import glob
import os

import numpy as np
import pytest

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "C*.npz")))
FLINK = {("poisson", "log"): 1, ("binomial", "logit"): 3, ("gaussian", "identity"): 7}


def load(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    g["family"], g["link"] = str(g["family"]), str(g["link"])
    return g


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_reference_outputs(path, oracle):
    g = load(path)
    fl = FLINK[(g["family"], g["link"])]
    X, Z, y, U, L = g["X"], g["Z"], g["y"], g["U"], g["L"]
    for k in range(4):
        got = oracle.loglik_faithful(X, Z, U, y, g["betas"][:, k], g["sigmas"][k], fl)
        assert got == g["loglik"][k]                                   # same arithmetic, same order: bit-exact
    for k in range(3):
        for faithful in (True, False):
            got = oracle.mvn_loglik(g["cov"], g["data"], g["eff_range"], g["thetas"][:, k], U, faithful=faithful)
            assert got == g["mvn_ll"][k]
        assert abs(oracle.logdet(g["cov"], g["data"], g["eff_range"], g["thetas"][:, k]) - g["logdet"][k]) <= 1e-13 * max(1, abs(g["logdet"][k]))
    nr = oracle.mcnr(X, Z, U, y, g["beta"], 1.0, fl)
    assert np.max(np.abs(g["beta"] + nr["beta_incr"] - g["mcnr_beta"])) <= 1e-12 * max(1.0, np.max(np.abs(g["mcnr_beta"])))
    assert abs(nr["sigma"] - g["mcnr_sigma"]) <= 1e-15
    ZL = Z @ L
    xb = X @ g["beta"]
    # X beta / Z L through numpy's BLAS instead of the reference's loops: rounding-level differences allowed here
    for k in range(3):
        assert abs(oracle.log_prob(ZL, xb, y, 0.9, fl, g["V"][:, k]) - g["log_prob"][k]) <= 1e-12 * abs(g["log_prob"][k])
        assert np.max(np.abs(oracle.log_grad(ZL, xb, y, 0.9, fl, g["V"][:, k]) - g["log_grad"][:, k])) <= 1e-12 * max(1.0, np.max(np.abs(g["log_grad"][:, k])))
    wu, ns, lam, ms, ta, seed, chain = g["hmc_settings"]
    ch = oracle.hmc_chain(ZL, L, xb, y, 1.0, fl, int(wu), int(ns), float(lam), int(ms), float(ta), int(seed), chain=int(chain))
    assert np.max(np.abs(ch["u"] - g["hmc_u"])) <= 1e-9 * max(1.0, np.max(np.abs(g["hmc_u"])))
    assert ch["accept"] == float(g["hmc_accept"]) and ch["steps"] == int(g["hmc_steps"])
    assert abs(ch["eps"] - float(g["hmc_eps"])) <= 1e-12
    # the three objective functors of likelihood.h at (beta, theta)
    Lobj = -oracle.loglik_faithful(X, Z, U, y, g["beta"], 1.0, fl)
    Dobj = -oracle.mvn_loglik(g["cov"], g["data"], g["eff_range"], g["theta"], U)
    assert Lobj == g["objectives"][0] and Dobj == g["objectives"][1]
    assert abs((Lobj + Dobj) - g["objectives"][2]) <= 1e-13 * abs(g["objectives"][2])


def test_family_terms_table(oracle):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "family_terms.npz"))
    for fl, yv, eta, sg, want in z["table"]:
        assert oracle.family_ll(yv, eta, sg, int(fl)) == want
    for k, want in z["log_factorial"]:
        assert oracle.log_factorial_approx(k) == want


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_reproduces_reference_outputs(path, gctx):
    import glmmrmcml_b200 as g_
    g = load(path)
    fam, link = g["family"], g["link"]
    X, Z, y, U, L = g["X"], g["Z"], g["y"], g["U"], g["L"]
    mdl = g_.Model(gctx, X, Z, y, fam, link)
    cv = g_.Covariance(gctx, g["cov"], g["data"], g["eff_range"])
    mdl.set_u(U)
    got = mdl.log_likelihood_batch(g["betas"], g["sigmas"])
    assert np.max(np.abs(got - g["loglik"]) / np.abs(g["loglik"])) <= 1e-10
    for k in range(3):
        v = cv.loglik(g["thetas"][:, k], U)
        assert abs(v - g["mvn_ll"][k]) <= 1e-10 * abs(g["mvn_ll"][k])
        assert abs(cv.logdet(g["thetas"][:, k]) - g["logdet"][k]) <= 1e-10 * max(1.0, abs(g["logdet"][k]))
    nr = mdl.mcnr(g["beta"], 1.0)
    assert np.max(np.abs(g["beta"] + nr["beta_incr"] - g["mcnr_beta"])) <= 1e-8 * max(1.0, np.max(np.abs(g["mcnr_beta"])))
    assert abs(nr["sigma"] - g["mcnr_sigma"]) <= 1e-10 * g["mcnr_sigma"]
    lp, G = mdl.log_prob_grad(L, g["beta"], 0.9, g["V"])
    assert np.max(np.abs(lp - g["log_prob"]) / np.abs(g["log_prob"])) <= 1e-10
    assert np.max(np.abs(G - g["log_grad"])) <= 1e-10 * max(1.0, np.max(np.abs(g["log_grad"])))
    wu, ns, lam, ms, ta, seed, chain = g["hmc_settings"]
    for variant in (1, 2, 3):
        g_.hmc_set_variant(variant)
        try:
            out = mdl.hmc_sample(L, g["beta"], 1.0, warmup=int(wu), nsamp_per_chain=int(ns), lam=float(lam), max_steps=int(ms),
                                 target_accept=float(ta), n_chains=1, chain_offset=int(chain), seed=int(seed))
        except g_.GmbError as e:
            if variant in (2, 3) and "does not fit" in str(e):
                continue
            raise
        finally:
            g_.hmc_set_variant(0)
        assert np.max(np.abs(out["u"] - g["hmc_u"])) <= 1e-7 * max(1.0, np.max(np.abs(g["hmc_u"])))
        assert abs(out["stats"]["accept_rate"] - float(g["hmc_accept"])) < 1e-12
        assert abs(out["stats"]["step_size_mean"] - float(g["hmc_eps"])) <= 1e-9
    # reference-named entry points on the same inputs
    assert abs(g_.mvn_ll(g["cov"], g["data"], g["eff_range"], g["theta"], U) - g["mvn_ll"][0]) <= 1e-10 * abs(g["mvn_ll"][0])
    bp = np.concatenate([g["beta"], [1.0]]) if fam == "gaussian" else g["beta"]
    aic = g_.aic_mcml(g["cov"], g["data"], g["eff_range"], Z, X, y, U, fam, link, bp, g["theta"])
    want = -2 * (-(g["objectives"][0]) - g["objectives"][1]) + 2 * (bp.size + g["theta"].size)
    assert abs(aic - want) <= 1e-10 * abs(want)
    opt = g_.mcml_optim(g["cov"], g["data"], g["eff_range"], Z, X, y, U, fam, link, np.concatenate([g["beta"], g["theta"], [1.0]]), 0, True)
    assert np.max(np.abs(opt["beta"] - g["mcnr_beta"])) <= 1e-8 * max(1.0, np.max(np.abs(g["mcnr_beta"])))
    mdl.close(); cv.close()
