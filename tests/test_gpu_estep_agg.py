"""E-step on the distinct rows of [X | Z] (estep.cu *_agg kernels, model.cu gmb_model_build_zd): observations that share
their row share eta, so zd = Z u is formed for the distinct rows only and the log-likelihood / MCNR sums run on those rows with the sufficient
statistics of their observations.  Checked against the per-observation kernels of the same library (1e-12) and against the oracle (1e-10; 1e-5
in fp32 mode), for the three families, single and batched evaluations, and the fall-backs."""
import numpy as np
import pytest

from glmmrmcml_b200 import synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def design(family, m=1500, nind=10):
    """config C2's cluster-period design (500 observations, 50 distinct rows) with a response of the given family"""
    cfg = synth.config2(m=m, nind=nind)
    rng = np.random.default_rng(21)
    eta = cfg["X"] @ cfg["beta"] + cfg["Z"] @ cfg["U"][:, 0]
    if family == "poisson":
        cfg.update(family="poisson", link="log", y=rng.poisson(np.exp(np.clip(eta, -3, 2.5))).astype(float))
    elif family == "gaussian":
        cfg.update(family="gaussian", link="identity", y=eta + 0.7 * rng.standard_normal(eta.size))
    return cfg


def models(g, gctx, cfg, precision="fp64"):
    """(aggregated, per-observation) models on the same samples"""
    out = []
    for agg in (True, False):
        mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"], precision=precision)
        g.estep_set_row_aggregation(agg)
        try:
            mdl.set_u(cfg["U"])
        finally:
            g.estep_set_row_aggregation(True)
        out.append(mdl)
    return out


@pytest.mark.parametrize("family", ["binomial", "poisson", "gaussian"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_aggregated_estep_equals_per_observation_estep_and_oracle(gctx, oracle, family, precision):
    import glmmrmcml_b200 as g
    cfg = design(family)
    fl = oracle.flink(cfg["family"], cfg["link"])
    ma, md = models(g, gctx, cfg, precision)
    assert ma.estep_rows() == 50 and md.estep_rows() == cfg["n"] == 500
    tol_lib, tol_orc = (1e-12, 1e-10) if precision == "fp64" else (2e-6, 1e-5)
    rng = np.random.default_rng(8)
    for trial in range(3):
        beta = cfg["beta"] + 0.1 * trial * rng.standard_normal(cfg["P"])
        sigma = 1.0 + 0.4 * trial
        want = oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], beta, sigma, fl)
        for rowstats in (True, False):                         # poisson / gaussian: row statistics (expanded per observation) and the stream
            g.estep_set_rowstats(rowstats)
            try:
                a, d = ma.log_likelihood(beta, sigma), md.log_likelihood(beta, sigma)
            finally:
                g.estep_set_rowstats(True)
            assert abs(a - d) <= tol_lib * abs(d), (family, rowstats, a, d)
            assert abs(a - want) <= tol_orc * abs(want), (family, rowstats, a, want)
        nra, nrd = ma.mcnr(beta, sigma), md.mcnr(beta, sigma)
        ref = oracle.mcnr(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], beta, sigma, fl)
        sc = np.max(np.abs(ref["xtwx"]))
        for key in ("xtwx", "score"):
            assert np.max(np.abs(nra[key] - nrd[key])) <= tol_lib * sc, (family, key)
            assert np.max(np.abs(nra[key] - ref[key])) <= tol_orc * sc, (family, key)
        assert abs(nra["sigma"] - nrd["sigma"]) <= tol_lib * nrd["sigma"] and abs(nra["sigma"] - ref["sigma"]) <= tol_orc * ref["sigma"]
        assert rel(nra["beta_incr"], nrd["beta_incr"]) <= (1e-9 if precision == "fp64" else 1e-4)
    for k in (3, 19, 300):                                     # batches: one pair of launches for the whole batch (binomial), row statistics otherwise
        B = np.asfortranarray(cfg["beta"][:, None] + 0.05 * rng.standard_normal((cfg["P"], k)))
        sg = np.full(k, 1.3)
        a, d = ma.log_likelihood_batch(B, sg), md.log_likelihood_batch(B, sg)
        assert rel(a, d) <= tol_lib and np.array_equal(a, ma.log_likelihood_batch(B, sg))
        assert abs(a[k - 1] - ma.log_likelihood(B[:, k - 1], 1.3)) <= 1e-13 * abs(a[k - 1])
    ma.close(); md.close()


def test_aggregation_falls_back_where_it_does_not_apply(gctx, oracle):
    import glmmrmcml_b200 as g
    # (a) binomial responses that are not all 0/1: moremaths.h:47-53 drops those observations from the log-likelihood but not from the MCNR sums
    cfg = design("binomial", m=300)
    y = cfg["y"].copy(); y[7] = 0.5
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], y, cfg["family"], cfg["link"])
    mdl.set_u(cfg["U"])
    assert mdl.estep_rows() == 500
    fl = oracle.flink(cfg["family"], cfg["link"])
    want = oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], y, cfg["beta"], 1.0, fl)
    assert abs(mdl.log_likelihood(cfg["beta"], 1.0) - want) <= 1e-10 * abs(want)
    mdl.close()
    # (b) fewer than 4 observations per distinct row: not worth it
    cfg = design("binomial", m=300, nind=3)
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    mdl.set_u(cfg["U"])
    assert mdl.estep_rows() == cfg["n"] == 150
    mdl.close()
    # (c) the sampler's row view is switched off after the samples were set: zd is rebuilt per observation, results unchanged
    cfg = design("poisson", m=400)
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    mdl.set_u(cfg["U"])
    before = mdl.log_likelihood(cfg["beta"], 1.0); nr0 = mdl.mcnr(cfg["beta"], 1.0)
    assert mdl.estep_rows() == 50
    L = synth.dense_chol_D(cfg["cov"], cfg["data"], cfg["theta"])
    g.hmc_set_row_aggregation(False)
    try:
        mdl.hmc_sample(L, cfg["beta"], 1.0, warmup=5, nsamp_per_chain=1, lam=1.0, max_steps=5, target_accept=0.8, n_chains=8, seed=1, keep_on_device=False, want_u=False)
        assert mdl.estep_rows() == 500
        after = mdl.log_likelihood(cfg["beta"], 1.0); nr1 = mdl.mcnr(cfg["beta"], 1.0)
    finally:
        g.hmc_set_row_aggregation(True)
    assert abs(after - before) <= 1e-12 * abs(before) and np.max(np.abs(nr1["xtwx"] - nr0["xtwx"])) <= 1e-12 * np.max(np.abs(nr0["xtwx"]))
    mdl.close()
