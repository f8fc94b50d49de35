"""Independent closed-form checks of the CPU oracle: scipy.stats densities with the reference's
constants substituted, central differences, a dense numpy transcription of the MCNR step, and the exact Gaussian
posterior for the sampler."""
import numpy as np
import pytest
from scipy import stats

from glmmrmcml_b200 import synth

PI_REF = 3.141593          # moremaths.h:21,76


def test_family_terms_against_scipy(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        eta = float(rng.normal() * 2); sg = float(rng.uniform(0.3, 3)); yb = float(rng.integers(0, 2)); yp = int(rng.integers(0, 40)); yg = float(rng.normal())
        p = 1 / (1 + np.exp(-eta))
        assert abs(oracle.family_ll(yb, eta, sg, 3) - stats.bernoulli.logpmf(int(yb), p)) <= 1e-12 * max(1.0, abs(eta))
        # poisson: exact pmf with log y! replaced by the reference's Ramanujan approximation
        lf = oracle.log_factorial_approx(yp)
        want = yp * eta - np.exp(eta) - lf
        assert abs(oracle.family_ll(yp, eta, sg, 1) - want) <= 1e-12 * max(1.0, abs(want))
        # gaussian with pi = 3.141593
        want = stats.norm.logpdf(yg, eta, sg) + 0.5 * np.log(2 * np.pi) - 0.5 * np.log(2 * PI_REF)
        assert abs(oracle.family_ll(yg, eta, sg, 7) - want) <= 1e-12 * max(1.0, abs(want))
    # Ramanujan's approximation is close to lgamma(n + 1) once pi is the truncated constant
    from scipy.special import gammaln
    for k in range(1, 80):
        assert abs(oracle.log_factorial_approx(k) - gammaln(k + 1)) < 2e-3
    assert oracle.log_factorial_approx(0) == 0.0


@pytest.mark.parametrize("make", [lambda: synth.config1(m=40), lambda: synth.config2(m=40), lambda: synth.config3(nloc=60, m=25),
                                  lambda: synth.config4(ncl=15, nt=6, k=1, m=30)], ids=["C1", "C2", "C3", "C4"])
def test_mvn_loglik_against_scipy(make, oracle):
    cfg = make()
    D = oracle.genD(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], chol=False)
    L = oracle.genD(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], chol=True)
    assert np.allclose(L @ L.T, D, rtol=1e-12, atol=1e-14)
    assert np.allclose(D, synth.dense_chol_D(cfg["cov"], cfg["data"], cfg["theta"]) @ synth.dense_chol_D(cfg["cov"], cfg["data"], cfg["theta"]).T, rtol=1e-10, atol=1e-13)
    want = np.mean(stats.multivariate_normal.logpdf(cfg["U"].T, mean=np.zeros(cfg["Q"]), cov=D, allow_singular=False))
    got = oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], cfg["U"])
    assert abs(got - want) <= 1e-9 * abs(want)
    sign, ld = np.linalg.slogdet(D)
    assert abs(oracle.logdet(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"]) - ld) <= 1e-9 * max(1.0, abs(ld))


def test_not_positive_definite_is_reported(oracle):
    cfg = synth.config2(m=3)
    with pytest.raises(np.linalg.LinAlgError):
        oracle.genD(cfg["cov"], cfg["data"], cfg["eff_range"], np.array([0.25, 1.7]), chol=True)
    assert np.isnan(oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], np.array([0.25, 1.7]), cfg["U"]))


@pytest.mark.parametrize("make", [lambda: synth.config1(m=4), lambda: synth.config3(nloc=30, m=4), lambda: synth.config4(ncl=10, nt=5, k=2, m=4)],
                         ids=["binomial", "gaussian", "poisson"])
def test_log_grad_is_the_gradient_of_log_prob(make, oracle):
    cfg = make()
    fl = oracle.flink(cfg["family"], cfg["link"])
    ZL = cfg["Z"] @ cfg["L"]; xb = cfg["X"] @ cfg["beta"]
    rng = np.random.default_rng(2)
    v = 0.5 * rng.standard_normal(cfg["Q"])
    g = oracle.log_grad(ZL, xb, cfg["y"], 0.8, fl, v)
    h = 1e-6
    for q in rng.choice(cfg["Q"], size=min(8, cfg["Q"]), replace=False):
        e = np.zeros(cfg["Q"]); e[q] = h
        fd = (oracle.log_prob(ZL, xb, cfg["y"], 0.8, fl, v + e) - oracle.log_prob(ZL, xb, cfg["y"], 0.8, fl, v - e)) / (2 * h)
        assert abs(fd - g[q]) <= 1e-6 * max(1.0, abs(g[q]))


@pytest.mark.parametrize("make", [lambda: synth.config2(m=30), lambda: synth.config3(nloc=25, m=12), lambda: synth.config4(ncl=8, nt=5, k=2, m=15)],
                         ids=["binomial", "gaussian", "poisson"])
def test_mcnr_against_dense_numpy_transcription(make, oracle):
    """mcmloptim.h:198-236 written with dense matrices exactly as the reference does (W as an n x n diagonal matrix)."""
    cfg = make()
    fl = oracle.flink(cfg["family"], cfg["link"])
    X, Z, U, y, beta = cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"]
    n, P = X.shape; m = U.shape[1]; sigma = 0.9
    zd = Z @ U; xb = X @ beta
    XtWX = np.zeros((P, P)); Wu = np.zeros((n, m)); sig = np.zeros(m)
    for j in range(m):
        eta = xb + zd[:, j]
        if cfg["family"] == "binomial":
            p = np.exp(eta) / (1 + np.exp(eta)); w = 1 / (p * (1 - p)); mu = p; dmu = 1 / (p * (1 - p)); phi = 1.0
        elif cfg["family"] == "poisson":
            w = np.exp(-eta); mu = np.exp(eta); dmu = np.exp(-eta); phi = 1.0
        else:
            w = np.ones(n); mu = eta; dmu = np.ones(n); phi = sigma ** 2
        W = np.diag(1 / (w * phi))
        resid = y - mu
        sig[j] = np.sqrt(np.sum((resid - resid.mean()) ** 2) / (n - 1))
        XtWX += X.T @ W @ X / m
        Wu[:, j] = np.diag(W) * dmu * resid
    incr = np.linalg.inv(XtWX) @ X.T @ Wu.mean(axis=1)
    got = oracle.mcnr(X, Z, U, y, beta, sigma, fl)
    assert np.allclose(got["xtwx"], XtWX, rtol=1e-12)
    assert np.allclose(got["beta_incr"], incr, rtol=1e-9, atol=1e-13)
    assert abs(got["sigma"] - sig.mean()) <= 1e-13


def test_oracle_chain_samples_the_exact_gaussian_posterior(oracle):
    """Gaussian-identity: v | y ~ N(A^-1 ZL^T (y - xb)/s^2, A^-1), A = I + ZL^T ZL / s^2."""
    cfg = synth.config3(nloc=12, m=4)
    sigma = 0.8
    ZL = cfg["Z"] @ cfg["L"]; xb = cfg["X"] @ cfg["beta"]
    A = np.eye(cfg["Q"]) + ZL.T @ ZL / sigma ** 2
    cov = np.linalg.inv(A); mean = cov @ ZL.T @ (cfg["y"] - xb) / sigma ** 2
    means = []
    allv = []
    for c in range(24):
        ch = oracle.hmc_chain(ZL, cfg["L"], xb, cfg["y"], sigma, 7, 150, 120, 1.5, 40, 0.9, 2024, chain=c, want_u=False)
        means.append(ch["v"][:, 1:].mean(axis=1)); allv.append(ch["v"][:, 1:])
    means = np.array(means)
    est = means.mean(axis=0); se = means.std(axis=0, ddof=1) / np.sqrt(len(means))
    assert np.all(np.abs(est - mean) <= 5 * se + 1e-12)
    var = np.concatenate(allv, axis=1).var(axis=1, ddof=1)
    assert np.all(np.abs(var / np.diag(cov) - 1) < 0.25)


def test_rng_streams(oracle):
    z = oracle.rng_normal_vec(123, 4, 5, 2, 200_001)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert np.array_equal(z[:100], oracle.rng_normal_vec(123, 4, 5, 2, 100))           # counter-based: prefix-stable
    assert not np.array_equal(z[:100], oracle.rng_normal_vec(123, 4, 6, 2, 100))
    u = np.array([oracle.rng_uniform(9, t, 0, 3) for t in range(2000)])
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.03


def test_digamma_against_scipy(oracle):
    """The oracle's stand-in for boost::math::digamma (beta family gradient, mcmlmodel.h:271)."""
    from scipy.special import digamma
    x = np.concatenate([np.linspace(0.05, 6, 60), np.linspace(6, 200, 40)])
    got = np.array([oracle.digamma(t) for t in x])
    assert np.max(np.abs(got - digamma(x)) / np.maximum(1.0, np.abs(digamma(x)))) <= 1e-10
