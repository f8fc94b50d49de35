"""The CPU oracle (oracle/oracle.cpp) against oracle/_ref — the reference's own headers compiled unmodified against
oracle/shim — on fresh random inputs, including edge cases the golden files do not carry."""
import numpy as np
import pytest

from glmmrmcml_b200 import synth

ref = pytest.importorskip("oracle.ref")
pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref.so not built (needs /root/reference)")

CASES = {
    "C1": lambda: synth.config1(m=33, seed=5),
    "C2": lambda: synth.config2(m=17, seed=6),
    "C3": lambda: synth.config3(nloc=31, m=9, seed=7),
    "C4": lambda: synth.config4(ncl=9, nt=4, k=3, m=21, seed=8),
    "single_column": lambda: synth.config2(m=1, seed=9),
}


@pytest.mark.parametrize("name", list(CASES))
def test_estep_and_cov(name, oracle):
    cfg = CASES[name]()
    fam, link = cfg["family"], cfg["link"]
    fl = oracle.flink(fam, link)
    X, Z, y, U = cfg["X"], cfg["Z"], cfg["y"], cfg["U"]
    for sig in (1.0, 0.4):
        assert oracle.loglik_faithful(X, Z, U, y, cfg["beta"], sig, fl) == ref.loglik(X, Z, U, y, cfg["beta"], sig, fam, link)
    for th in (cfg["theta"], cfg["theta"] * np.array([1.4, 0.9])[: cfg["theta"].size]):
        want = ref.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], th, U)
        assert oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], th, U, faithful=True) == want
        assert oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], th, U, faithful=False) == want
    if U.shape[1] > 1:
        start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
        b, s = ref.mcnr(cfg["cov"], cfg["data"], cfg["eff_range"], X, Z, U, y, fam, link, start)
        o = oracle.mcnr(X, Z, U, y, cfg["beta"], 1.0, fl)
        assert np.max(np.abs(cfg["beta"] + o["beta_incr"] - b)) <= 1e-12 * max(1.0, np.max(np.abs(b)))
        assert o["sigma"] == s
        of = oracle.mcnr(X, Z, U, y, cfg["beta"], 1.0, fl, faithful=True)
        assert np.array_equal(of["xtwx"], o["xtwx"]) and np.array_equal(of["score"], o["score"])


@pytest.mark.parametrize("name", ["C1", "C3", "C4"])
def test_sampler_chain_is_bitwise_the_reference_chain(name, oracle):
    """oracle.hmc_chain and the reference's mcmcRunHMC::sample, fed the same Philox stream, visit the same states."""
    cfg = CASES[name]()
    fam, link = cfg["family"], cfg["link"]
    fl = oracle.flink(fam, link)
    X, Z, y, L = cfg["X"], cfg["Z"], cfg["y"], cfg["L"]
    # Z L and X beta as the reference forms them (plain loops): use the oracle's own GEMM, not numpy's BLAS
    ZL = oracle.gemm(Z, L)
    xb = np.zeros(cfg["n"]); oracle.lib().orc_xb(cfg["n"], cfg["P"], np.asfortranarray(X).ctypes.data_as(oracle._dp), cfg["beta"].ctypes.data_as(oracle._dp), xb.ctypes.data_as(oracle._dp))
    for lam, ms in ((0.05, 15), (5.0, 7)):
        u, st = ref.mcmc_sample(X, Z, L, y, cfg["beta"], fam, link, 20, 10, lam, 1.0, ms, 0.9, 4242, chain=1)
        ch = oracle.hmc_chain(ZL, L, xb, y, 1.0, fl, 20, 10, lam, ms, 0.9, 4242, chain=1)
        assert np.max(np.abs(ch["u"] - u)) <= 1e-12 * max(1.0, np.max(np.abs(u)))
        assert ch["accept"] == st["accept"] and ch["steps"] == st["steps"] and abs(ch["eps"] - st["eps"]) <= 1e-15


def test_family_terms_and_helpers(oracle):
    rng = np.random.default_rng(0)
    for fl in (1, 2, 3, 4, 5, 6, 7, 8):
        for _ in range(50):
            y = float(rng.integers(0, 2)) if fl in (3, 4, 5, 6) else (float(rng.integers(0, 30)) if fl in (1, 2) else float(rng.normal() + 2.5))
            mu = float(rng.uniform(0.05, 0.9)) if fl in (2, 5) else (float(-rng.uniform(0.05, 3)) if fl == 4 else float(rng.normal() * 2))
            sg = float(rng.uniform(0.3, 3))
            a, b = oracle.family_ll(y, mu, sg, fl), ref.family_ll(y, mu, sg, fl)
            assert a == b or (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-14 * abs(b)     # fl 6 uses erfc vs R::pnorm stand-in
    eta = rng.normal(size=20)
    for link, lnk in (("log", 0), ("identity", 1), ("logit", 2)):
        want = ref.detadmu(eta, link)
        got = np.array([1.0 if lnk == 1 else (np.exp(-1.0 * e) if lnk == 0 else 1 / ((np.exp(e) / (1 + np.exp(e))) * (1 - np.exp(e) / (1 + np.exp(e))))) for e in eta])
        assert np.allclose(got, want, rtol=1e-15, atol=0)
    A = rng.normal(size=(6, 6)); Lm = np.linalg.cholesky(A @ A.T + 6 * np.eye(6)); u = rng.normal(size=6)
    assert np.allclose(ref.forward_sub(Lm, u), np.linalg.solve(Lm, u), rtol=1e-13)


ALL_CODES = {
    1: ("poisson", "log"), 2: ("poisson", "identity"), 3: ("binomial", "logit"), 4: ("binomial", "log"), 5: ("binomial", "identity"),
    6: ("binomial", "probit"), 7: ("gaussian", "identity"), 8: ("gaussian", "log"),
    # the map's keys are lower case (mcmlmodel.h:83-86); R's family object says "Gamma", so 9-11 cannot be reached from R — the headers can
    9: ("gamma", "log"), 10: ("gamma", "inverse"), 11: ("gamma", "identity"), 12: ("beta", "logit"),
}


def _model_for_code(fl, rng, n=40, Q=6):
    """A small model whose linear predictor stays inside the code's domain (log of a mean, a probability, a positive mean)."""
    grp = np.arange(n) % Q
    Z = np.zeros((n, Q)); Z[np.arange(n), grp] = 1.0
    A = rng.normal(size=(Q, Q)); L = np.linalg.cholesky(A @ A.T / Q + np.eye(Q)) * 0.05
    X = np.column_stack([np.ones(n), rng.uniform(-1, 1, size=n)])
    if fl in (2, 10, 11):   beta = np.array([2.0, 0.3])        # eta > 0
    elif fl == 4:           beta = np.array([-1.5, 0.2])       # eta < 0
    elif fl in (5, 12):     beta = np.array([0.5, 0.1])        # 0 < eta < 1 (code 12's log-density takes eta as the mean, moremaths.h:98-99)
    else:                   beta = np.array([0.3, -0.4])
    if fl in (1, 2):        y = rng.poisson(3.0, size=n).astype(float)
    elif fl in (3, 4, 5, 6): y = rng.integers(0, 2, size=n).astype(float)
    elif fl == 7:           y = rng.normal(size=n) + 0.5
    elif fl == 8:           y = rng.uniform(3.0, 9.0, size=n)   # log(log y) must exist: the constructor and the density both take logs
    elif fl == 12:          y = rng.uniform(0.05, 0.95, size=n)
    else:                   y = rng.gamma(2.0, 1.0, size=n) + 0.05
    return X, Z, L, y, beta


@pytest.mark.parametrize("fl", sorted(ALL_CODES))
def test_log_prob_and_log_grad_of_every_family_link_code(fl, oracle):
    """mcmlmodel.h:138-153 and :156-279 for ALL twelve codes of the table at :74-87: the reference's own headers (libref) against the
    oracle's restatement — the GPU tests of codes 2, 4, 5, 6, 8 compare the kernels with this oracle."""
    fam, link = ALL_CODES[fl]
    assert oracle.flink(fam, link) == fl
    rng = np.random.default_rng(100 + fl)
    X, Z, L, y, beta = _model_for_code(fl, rng)
    ZL = oracle.gemm(Z, L); xb = X @ beta
    y_model = np.log(y) if fl == 8 else y                      # mcmlmodel.h:90-92
    for sig in (1.0, 1.7):
        for _ in range(3):
            v = rng.normal(size=Z.shape[1])
            lp_ref = ref.log_prob(X, Z, L, y, beta, sig, fam, link, v)
            lg_ref = ref.log_grad(X, Z, L, y, beta, sig, fam, link, v)
            assert np.isfinite(lp_ref) and np.all(np.isfinite(lg_ref)), (fl, lp_ref)
            lp = oracle.log_prob(ZL, xb, y_model, sig, fl, v)
            lg = oracle.log_grad(ZL, xb, y_model, sig, fl, v)
            assert abs(lp - lp_ref) <= 1e-12 * abs(lp_ref), (fl, lp, lp_ref)
            assert np.max(np.abs(lg - lg_ref)) <= 1e-12 * max(1.0, np.max(np.abs(lg_ref))), fl


def test_family_terms_of_the_gamma_and_beta_codes(oracle):
    """moremaths.h:83-99 (codes 9-12): tgamma / lgamma forms, bit for bit or to the last ulp."""
    rng = np.random.default_rng(12)
    for fl in (9, 10, 11, 12):
        for _ in range(50):
            y = float(rng.uniform(0.05, 0.95)) if fl == 12 else float(rng.gamma(2.0, 1.0) + 0.05)
            mu = float(rng.uniform(0.05, 0.95)) if fl == 12 else (float(rng.normal()) if fl == 9 else float(rng.uniform(0.2, 4)))
            sg = float(rng.uniform(0.3, 3))
            a, b = oracle.family_ll(y, mu, sg, fl), ref.family_ll(y, mu, sg, fl)
            assert np.isfinite(b) and (a == b or abs(a - b) <= 1e-15 * abs(b)), (fl, a, b)


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_one_iteration_of_the_oracle_mcml_loop_is_the_reference_iteration(name, oracle):
    """oracle/mcml_loop.py (the checker of tests/test_gpu_fit_parity.py) against the reference's own headers, piece by piece, for the first
    MCML iteration of src/mcml_full.cpp:83-126: the samples are mcmcRunHMC::sample's on the same Philox stream (Q x (m + 1)), the beta
    step is mcmloptim::mcnr on the first m columns (niter_ = m, App. B #1), and the theta step minimises the reference's D_likelihood over
    all m + 1 columns (checked as a local optimum of -MCMLDmatrix::loglik; rminqa's BOBYQA itself is not available)."""
    from oracle import mcml_loop
    cfg = CASES[name]()
    fam, link = cfg["family"], cfg["link"]
    X, Z, y = cfg["X"], cfg["Z"], cfg["y"]
    cov = (cfg["cov"], cfg["data"], cfg["eff_range"])
    P, R = cfg["P"], cfg["theta"].size
    start = np.concatenate([cfg["beta"] * 0.8, cfg["theta"] * 1.2, [1.0]])
    m, warm, lam, ms, tgt, seed = 24, 30, 0.5, 20, 0.9, 777
    fit = mcml_loop.mcml_full(*cov, Z, X, y, fam, link, start, mcnr=True, m=m, maxiter=1, warmup=warm, tol=1e-12, lam=lam, maxsteps=ms,
                              target_accept=tgt, seed=seed)
    assert fit["iter"] == 1 and fit["u"].shape == (cfg["Q"], m + 1)
    L0 = oracle.genD(*cov, start[P:P + R], chol=True)
    u_ref, _ = ref.mcmc_sample(X, Z, L0, y, start[:P], fam, link, warm, m, lam, 1.0, ms, tgt, (seed + mcml_loop.GOLDEN) & mcml_loop.MASK, chain=0)
    assert np.max(np.abs(fit["u"] - u_ref)) <= 1e-12 * max(1.0, np.max(np.abs(u_ref)))
    b_ref, _ = ref.mcnr(*cov, X, Z, u_ref[:, :m], y, fam, link, start)
    assert np.max(np.abs(fit["beta"] - b_ref)) <= 1e-10 * max(1.0, np.max(np.abs(b_ref)))
    f0 = -ref.mvn_loglik(*cov, fit["theta"], u_ref)
    for r in range(R):
        for h in (-1e-3, 1e-3):
            th = fit["theta"].copy(); th[r] = max(th[r] + h, 1e-6)
            assert -ref.mvn_loglik(*cov, th, u_ref) >= f0 - 1e-9 * abs(f0), (r, h)


def test_the_mcem_beta_step_of_the_oracle_loop_minimises_the_reference_objective(oracle):
    """method = 'mcem' (l_optim, mcmloptim.h:71-88): the oracle loop's beta after one iteration is a local minimum of the reference's own
    L_likelihood (-mcmlModel::log_likelihood on the first m columns of that iteration's samples)."""
    from oracle import mcml_loop
    cfg = CASES["C1"]()
    fam, link = cfg["family"], cfg["link"]
    X, Z, y = cfg["X"], cfg["Z"], cfg["y"]
    cov = (cfg["cov"], cfg["data"], cfg["eff_range"])
    P = cfg["P"]
    start = np.concatenate([cfg["beta"] * 0.8, cfg["theta"], [1.0]])
    m = 24
    fit = mcml_loop.mcml_full(*cov, Z, X, y, fam, link, start, mcnr=False, m=m, maxiter=1, warmup=30, tol=1e-12, lam=0.5, maxsteps=20,
                              target_accept=0.9, seed=99)
    U = np.asfortranarray(fit["u"][:, :m])
    f0 = -ref.loglik(X, Z, U, y, fit["beta"], 1.0, fam, link)
    for p in range(P):
        for h in (-1e-3, 1e-3):
            b = fit["beta"].copy(); b[p] += h
            assert -ref.loglik(X, Z, U, y, b, 1.0, fam, link) >= f0 - 1e-10 * abs(f0), (p, h)
