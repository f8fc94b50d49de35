"""GPU parity of the E-step kernels (K1 zd = Z u, K2 log-likelihood, K3 MCNR sums) against the CPU oracle.

Tolerance: 1e-10 relative in fp64 (BASELINE.json north_star).  All calls go through the C-ABI (ctypes).
"""
import numpy as np
import pytest

from glmmrmcml_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


CASES = {
    "C1": lambda: synth.config1(m=250),
    "C2": lambda: synth.config2(m=2000),
    "C3": lambda: synth.config3(nloc=250, m=250),
    "C4": lambda: synth.config4(ncl=100, nt=10, k=1, m=512),
    "C5": lambda: synth.config5(nloc=300, nobs=10, m=600),
    "ragged": lambda: synth.config4(ncl=37, nt=7, k=3, m=131),      # n = 777, odd sizes everywhere
}


# C1 and C2 repeat every row of [X | Z] ten times: by default their E-step runs on the 50 distinct rows (estep.cu: *_agg kernels); the
# "-dense" variants switch that off so that the per-observation kernels (factor matrix, TMA MCNR pass) stay covered on the same data
@pytest.fixture(scope="module", params=list(CASES) + ["C1-dense", "C2-dense"])
def case(request, gctx):
    import glmmrmcml_b200 as g
    name = request.param
    cfg = CASES[name.replace("-dense", "")]()
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    g.estep_set_row_aggregation(not name.endswith("-dense"))
    try:
        mdl.set_u(cfg["U"])
    finally:
        g.estep_set_row_aggregation(True)
    assert mdl.estep_rows() == (50 if name in ("C1", "C2") else cfg["n"])
    yield cfg, mdl
    mdl.close()


def test_loglik_matches_oracle(case, oracle):
    cfg, mdl = case
    fl = oracle.flink(cfg["family"], cfg["link"])
    rng = np.random.default_rng(7)
    for trial in range(3):
        beta = cfg["beta"] + 0.1 * trial * rng.standard_normal(cfg["P"])
        sigma = 1.0 + 0.3 * trial
        want = oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], beta, sigma, fl)
        got = mdl.log_likelihood(beta, sigma)
        assert abs(got - want) <= RTOL * abs(want), (got, want)


@pytest.mark.parametrize("k", [2, 8, 19, 256])
def test_loglik_multi_kernel_equals_single_kernel(gctx, oracle, k):
    """Batched binomial/logit evaluations run as one launch, 8 parameter vectors per pass over the factor matrix (estep.cu:
    loglik_logit_factor_multi_kernel); batches shorter than, equal to and not a multiple of 8, with a parameter vector large enough to
    push groups of factors past the overflow guard.  Same sums in another partition: 1e-13 relative against the single-launch kernel,
    1e-10 against the oracle, and bitwise reproducible."""
    import glmmrmcml_b200 as g
    cfg = synth.config2(m=1500)
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    g.estep_set_row_aggregation(False)             # the factor-matrix kernels are the per-observation path
    try:
        mdl.set_u(cfg["U"])
    finally:
        g.estep_set_row_aggregation(True)
    assert mdl.estep_rows() == cfg["n"]
    rng = np.random.default_rng(5)
    betas = np.asfortranarray(cfg["beta"][:, None] + 0.2 * rng.standard_normal((cfg["P"], k)))
    betas[:, k - 1] *= 150.0                       # |eta| ~ 100: products of 8 factors overflow, term-by-term path
    try:
        g.estep_set_multi(False); single = mdl.log_likelihood_batch(betas, np.ones(k))
    finally:
        g.estep_set_multi(True)
    multi = mdl.log_likelihood_batch(betas, np.ones(k))
    assert np.all(np.isfinite(multi)) and rel(multi, single) <= 1e-13 and np.max(np.abs(multi / single - 1)) <= 1e-13
    assert np.array_equal(multi, mdl.log_likelihood_batch(betas, np.ones(k)))
    fl = oracle.flink(cfg["family"], cfg["link"])
    for e in (0, k - 1):
        want = oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], betas[:, e], 1.0, fl)
        assert abs(multi[e] - want) <= RTOL * abs(want)
    mdl.close()


def test_loglik_batch_equals_single(case):
    cfg, mdl = case
    rng = np.random.default_rng(11)
    betas = cfg["beta"][:, None] + 0.05 * rng.standard_normal((cfg["P"], 9))
    sig = 1.0 + 0.1 * np.arange(9)
    b = mdl.log_likelihood_batch(betas, sig)
    s = np.array([mdl.log_likelihood(betas[:, k], sig[k]) for k in range(9)])
    if cfg["family"] == "binomial":      # batches of >= 8 take the one-launch kernel (another partition of the same sum)
        assert rel(b, s) <= 1e-13
    else:
        assert np.array_equal(b, s)      # same kernel, deterministic reduction -> bitwise equal


def test_loglik_is_deterministic(case):
    cfg, mdl = case
    a = mdl.log_likelihood(cfg["beta"], 1.0)
    for _ in range(3):
        assert mdl.log_likelihood(cfg["beta"], 1.0) == a


def test_niter_quirk(case, oracle):
    """niter_ < cols(u): the E-step averages the leading niter columns only (mcmlmodel.h:73,295)."""
    cfg, mdl = case
    fl = oracle.flink(cfg["family"], cfg["link"])
    m = cfg["U"].shape[1]
    mdl.set_u(cfg["U"], m_total=m, niter_total=m - 1)
    want = oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"], 1.0, fl, niter=m - 1)
    got = mdl.log_likelihood(cfg["beta"], 1.0)
    mdl.set_u(cfg["U"])
    assert abs(got - want) <= RTOL * abs(want)


def test_mcnr_matches_oracle(case, oracle):
    cfg, mdl = case
    fl = oracle.flink(cfg["family"], cfg["link"])
    for sigma in (1.0, 0.7):
        want = oracle.mcnr(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"], sigma, fl)
        got = mdl.mcnr(cfg["beta"], sigma)
        assert rel(got["xtwx"], want["xtwx"]) <= RTOL
        # the score is a sum of signed terms: scale by the size of its terms (max |X|^T mean |Wu|) rather than the result
        assert np.max(np.abs(got["score"] - want["score"])) <= RTOL * max(np.max(np.abs(want["score"])), np.max(np.abs(want["xtwx"])))
        assert abs(got["sigma"] - want["sigma"]) <= RTOL * abs(want["sigma"])
        assert rel(got["beta_incr"], want["beta_incr"]) <= 1e-8     # solve amplifies by cond(X^T W X)


def test_errors(gctx):
    import glmmrmcml_b200 as g
    cfg = synth.config1(m=8)
    with pytest.raises(g.GmbError) as e:
        g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "cloglog")
    assert e.value.code == 2
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
    with pytest.raises(g.GmbError) as e:
        mdl.log_likelihood(cfg["beta"])          # no samples yet
    assert e.value.code == 6
    mdl.close()


def test_rowstat_evaluation_equals_the_stream(gctx, oracle):
    """poisson / gaussian: the O(n) evaluation from row statistics of zd (default) and the streaming kernel give the same value to
    rounding, for several beta / sigma on the same sample matrix and with the niter quirk (fewer columns used than stored)."""
    import glmmrmcml_b200 as g
    for cfg, fam, link in ((synth.config4(ncl=40, nt=10, k=2, m=700), "poisson", "log"), (synth.config3(nloc=150, m=500), "gaussian", "identity")):
        mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], fam, link)
        fl = oracle.flink(fam, link)
        rng = np.random.default_rng(1)
        for niter in (None, cfg["U"].shape[1] - 7):
            mdl.set_u(cfg["U"], niter_total=niter)
            for k in range(3):
                beta = cfg["beta"] + 0.05 * rng.standard_normal(cfg["P"]); sg = 0.7 + 0.2 * k
                try:
                    g.estep_set_rowstats(False); a = mdl.log_likelihood(beta, sg)
                finally:
                    g.estep_set_rowstats(True)
                b = mdl.log_likelihood(beta, sg)
                want = oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], beta, sg, fl, niter=niter)
                assert abs(a - want) <= 1e-10 * abs(want)
                assert abs(b - want) <= 1e-10 * abs(want), (a, b, want)
        mdl.close()


@pytest.mark.parametrize("name", ["C4", "C5"])
def test_sparse_zd_build_equals_dense_contraction(gctx, oracle, name):
    """zd = Z u through the sparse form of Z (indicator designs, Q >= 64) against the dense DMMA contraction and the oracle: same sums
    without the zero terms."""
    import glmmrmcml_b200 as g
    cfg = CASES[name]()
    fl = oracle.flink(cfg["family"], cfg["link"])
    vals = {}
    try:
        for on in (True, False):
            g.estep_set_sparse_zd(on)
            mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
            mdl.set_u(cfg["U"])
            nr = mdl.mcnr(cfg["beta"], 1.0)
            vals[on] = (mdl.log_likelihood(cfg["beta"], 1.0), nr["xtwx"], nr["score"])
            mdl.close()
    finally:
        g.estep_set_sparse_zd(True)
    want = oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"], 1.0, fl)
    assert abs(vals[True][0] - want) <= RTOL * abs(want)
    assert abs(vals[True][0] - vals[False][0]) <= 1e-12 * abs(want)
    assert rel(vals[True][1], vals[False][1]) <= 1e-12 and rel(vals[True][2], vals[False][2]) <= 1e-10
