"""GPU tests of the reference-named entry points (src/mcml_optim.cpp, src/mcml_full.cpp) against optimisers and stencils
driven by the CPU oracle's objectives on the same fixed sample matrix.

The reference minimises with BOBYQA (rminqa, not available offline); any correct bounded optimiser reaches the same
optimum within the MCML tolerance, so the M-steps are compared with scipy minimisers of the ORACLE's
objectives, and the Hessian with the optimhess stencil (mcmloptim.h:333-355) evaluated on the oracle's objective."""
import numpy as np
import pytest
from scipy.optimize import minimize

from glmmrmcml_b200 import synth

pytestmark = pytest.mark.gpu


def small_rct(m=400, seed=11):
    return synth.config2(m=m, seed=seed, ncl=8, nt=4, nind=6)      # n = 192, P = 5, Q = 32, 8 blocks of 4


def oracle_objectives(cfg, oracle):
    fl = oracle.flink(cfg["family"], cfg["link"])
    zd = oracle.gemm(cfg["Z"], cfg["U"])
    def L_obj(beta, sigma=1.0):                                    # L_likelihood, likelihood.h:57-64
        return -oracle.loglik_zd(zd, cfg["X"] @ beta, cfg["y"], sigma, fl)
    def D_obj(theta):                                              # D_likelihood, likelihood.h:40-45
        if np.any(np.asarray(theta) < 1e-6):
            return 1e300
        v = -oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], np.asarray(theta, dtype=np.float64), cfg["U"])
        return v if np.isfinite(v) else 1e300                      # D(theta) not positive definite (ar1 parameter >= 1)
    return L_obj, D_obj


def test_mcml_optim_mcem_matches_scipy_on_oracle_objectives(gctx, oracle):
    """mcml_optim(mcnr = FALSE): l_optim then d_optim (src/mcml_optim.cpp:55-60)."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = small_rct()
    L_obj, D_obj = oracle_objectives(cfg, oracle)
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    fit = g.mcml_optim(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["U"], cfg["family"], cfg["link"], start, 0, False)
    rb = minimize(L_obj, cfg["beta"], method="BFGS", options=dict(gtol=1e-9))
    rt = minimize(D_obj, cfg["theta"], method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-14, maxiter=2000))
    assert np.max(np.abs(fit["beta"] - rb.x)) <= 2e-5, (fit["beta"], rb.x)
    assert np.max(np.abs(fit["theta"] - rt.x)) <= 2e-5, (fit["theta"], rt.x)
    # and the optimum is at least as good as scipy's on the oracle's own objective
    assert L_obj(fit["beta"]) <= rb.fun + 1e-9 * abs(rb.fun)
    assert D_obj(fit["theta"]) <= rt.fun + 1e-9 * abs(rt.fun)


def test_mcml_optim_gaussian_estimates_sigma(gctx, oracle):
    """gaussian family: l_optim runs over (beta, sigma) with sigma >= 0 (mcmloptim.h:79-83)."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = synth.config3(nloc=60, m=300)
    fl = oracle.flink("gaussian", "identity")
    zd = oracle.gemm(cfg["Z"], cfg["U"])
    obj = lambda p: -oracle.loglik_zd(zd, cfg["X"] @ p[:1], cfg["y"], p[1], fl) if p[1] > 0 else np.inf
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.3]])
    fit = g.mcml_optim(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["U"], "gaussian", "identity", start, 0, False)
    r = minimize(obj, [cfg["beta"][0], 1.3], method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=4000))
    assert abs(fit["beta"][0] - r.x[0]) <= 1e-4 and abs(fit["sigma"] - r.x[1]) <= 1e-4, (fit, r.x)


def test_mcml_hess_is_the_optimhess_stencil_of_the_oracle_objective(gctx, oracle):
    """mcml_hess (src/mcml_optim.cpp:263-285): FD Hessian of F_likelihood(importance = false) with step tol.  Compared with the
    same stencil on the oracle's objective at a larger step, where finite differences resolve the curvature (SURVEY §7 hard parts:
    at the reference's default 1e-5 the stencil amplifies 1e-13 differences in ll to 1e-3 in H)."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = small_rct(m=300)
    L_obj, D_obj = oracle_objectives(cfg, oracle)
    P = cfg["P"]
    F = lambda X: np.array([L_obj(X[:P, k]) + D_obj(X[P:, k]) for k in range(X.shape[1])])
    x0 = np.concatenate([cfg["beta"], cfg["theta"]])
    tol = 1e-3
    H = g.mcml_hess(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["U"], cfg["family"], cfg["link"], x0, tol, 0)
    lower = np.concatenate([np.full(P, -np.inf), np.full(2, 1e-6)])
    Hw, nfev = g.fd_hessian(F, x0, tol, lower=lower, upper=np.full(P + 2, np.inf), usebounds=True)
    assert H.shape == (P + 2, P + 2) and np.allclose(H, H.T)
    assert np.max(np.abs(H - Hw)) <= 1e-5 * np.max(np.abs(Hw)), np.max(np.abs(H - Hw)) / np.max(np.abs(Hw))
    # the beta block is X^T diag(mean_j w_j) X (analytic), which the stencil must reproduce to FD accuracy
    nr = oracle.mcnr(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"], 1.0, oracle.flink(cfg["family"], cfg["link"]))
    assert np.max(np.abs(H[:P, :P] - nr["xtwx"])) <= 1e-4 * np.max(np.abs(nr["xtwx"]))


def test_mcml_simlik_matches_scipy_on_oracle_objective(gctx, oracle):
    """mcml_simlik: joint BOBYQA over (beta, theta) of F_likelihood with importance weights (likelihood.h:88-108), evaluated in
    log space; the constant denominator does not move the optimum."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = small_rct(m=300)
    L_obj, D_obj = oracle_objectives(cfg, oracle)
    P = cfg["P"]
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    fit = g.mcml_simlik(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["U"], cfg["family"], cfg["link"], start, 0)
    F = lambda p: L_obj(p[:P]) + D_obj(p[P:])
    r = minimize(F, start[:P + 2], method="BFGS", options=dict(gtol=1e-8))
    got = np.concatenate([fit["beta"], fit["theta"]])
    assert np.max(np.abs(got - r.x)) <= 5e-5, (got, r.x)


def test_mcml_full_converges_and_agrees_across_seeds(gctx, oracle):
    """mcml_full (src/mcml_full.cpp:41-148) with the native sampler: converges, returns Q x (m + 1) samples, and two seeds agree within
    a few Monte-Carlo tolerances (the reference's own stopping rule is max |delta| < tol between iterations)."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = small_rct(m=8)
    start = np.concatenate([cfg["beta"] * 0.5, [0.4, 0.5], [1.0]])
    fits = []
    for seed in (1, 2):
        fit = g.mcml_full(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start,
                          mcnr=True, m=2000, maxiter=40, warmup=200, tol=2e-2, verbose=False, lam=1.0, maxsteps=50, target_accept=0.9,
                          n_chains=100, seed=seed)
        assert fit["u"].shape == (cfg["Q"], 2001) and np.all(np.isfinite(fit["u"]))
        assert np.all(np.isfinite(fit["beta"])) and np.all(fit["theta"] > 0)
        assert fit["iter"] >= 2
        fits.append(fit)
    assert np.max(np.abs(fits[0]["beta"] - fits[1]["beta"])) <= 0.1
    assert np.max(np.abs(fits[0]["theta"] - fits[1]["theta"])) <= 0.1
    # the MCNR fixed point solves the Monte-Carlo score equation: one more oracle MCNR step on the returned samples barely moves beta
    fl = oracle.flink(cfg["family"], cfg["link"])
    nr = oracle.mcnr(cfg["X"], cfg["Z"], np.asfortranarray(fits[0]["u"][:, :2000]), cfg["y"], fits[0]["beta"], 1.0, fl)
    assert np.max(np.abs(nr["beta_incr"])) <= 5e-2


def test_model_mcml_class_runs_the_reference_call_sequence(gctx):
    """ModelMCML$MCML(y, usestan = FALSE): mcml_full -> mcml_hess -> aic_mcml (R/R6ModelExtMCML.R:399-553)."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = small_rct(m=8)
    mod = g.ModelMCML(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["family"], cfg["link"], cfg["beta"], cfg["theta"])
    mod.mcmc_options.update(warmup=100, samps=500, lam=1.0, maxsteps=30)
    out = mod.MCML(cfg["y"], verbose=False, tol=5e-2, max_iter=10, method="mcnr", n_chains=50, seed=3)
    k = cfg["P"] + 2
    assert out["hessian"].shape == (k, k) and np.isfinite(out["aic"])
    assert out["u"].shape == (cfg["Q"], 501)


def test_model_mcml_usestan_branch_runs_on_the_native_sampler(gctx):
    """ModelMCML$MCML(y, usestan = TRUE) (R/R6ModelExtMCML.R:238-373): the R-level loop — sample, mcml_optim, refresh L from the previous
    iterate — with the Stan call replaced by the native sampler, then the simulated-likelihood step; the estimates agree with the
    usestan = FALSE path (mcml_full) within Monte-Carlo error and the sample matrix has `samps` columns like Stan's draws."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = small_rct(m=8)
    mod = g.ModelMCML(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["family"], cfg["link"], cfg["beta"], cfg["theta"])
    mod.mcmc_options.update(warmup=100, samps=2000, lam=1.0, maxsteps=30)
    a = mod.MCML(cfg["y"], verbose=False, tol=5e-2, max_iter=10, method="mcnr", usestan=True, sim_lik_step=True, n_chains=50, seed=3)
    b = mod.MCML(cfg["y"], verbose=False, tol=5e-2, max_iter=10, method="mcnr", usestan=False, n_chains=50, seed=3)
    assert a["u"].shape == (cfg["Q"], 2000) and b["u"].shape == (cfg["Q"], 2001)
    assert a["iter"] >= 2 and np.all(np.isfinite(a["beta"])) and np.all(a["theta"] > 0)
    assert np.max(np.abs(a["beta"] - b["beta"])) <= 0.25
    assert a["hessian"].shape == (cfg["P"] + 2, cfg["P"] + 2) and np.isfinite(a["aic"])
    L = mod.chol_D(cfg["theta"])
    assert np.allclose(L, synth.dense_chol_D(cfg["cov"], cfg["data"], cfg["theta"]), rtol=1e-10, atol=1e-12)


def test_model_mcml_gaussian_takes_hessian_ses(gctx):
    """Gaussian family through ModelMCML$MCML with se_theta = TRUE: mcml_hess receives (beta, theta, sigma) like `start = theta` in
    R/R6ModelExtMCML.R:464-474 (mcmloptim ctor needs P + R + 1 values for gaussian models, mcmloptim.h:30)."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = synth.config3(nloc=40, m=8)
    mod = g.ModelMCML(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], "gaussian", "identity", cfg["beta"], cfg["theta"], var_par=1.0)
    mod.mcmc_options.update(warmup=100, samps=400, lam=1.0, maxsteps=30)
    out = mod.MCML(cfg["y"], verbose=False, tol=5e-2, max_iter=4, method="mcem", n_chains=40, seed=5)
    k = cfg["P"] + 2
    assert out["hessian"].shape == (k, k) and np.all(np.isfinite(out["hessian"])) and np.isfinite(out["aic"])
    assert out["sigma"] > 0


def test_mcml_hess_with_many_fixed_effects(gctx, oracle):
    """P = 25: the optimhess stencil has 4 (P + R)^2 = 2916 points, more parameter values than one staging buffer holds — the batch is
    chunked inside the library (the reference has no limit, mcmloptim.h:333-355)."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = small_rct(m=64)
    rng = np.random.default_rng(3)
    Xw = np.asfortranarray(np.column_stack([cfg["X"], 0.3 * rng.standard_normal((cfg["n"], 25 - cfg["P"]))]))
    beta = np.concatenate([cfg["beta"], 0.1 * rng.standard_normal(25 - cfg["P"])])
    x0 = np.concatenate([beta, cfg["theta"]])
    H = g.mcml_hess(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], Xw, cfg["y"], cfg["U"], cfg["family"], cfg["link"], x0, 1e-3, 0)
    assert H.shape == (27, 27) and np.all(np.isfinite(H))
    nr = oracle.mcnr(Xw, cfg["Z"], cfg["U"], cfg["y"], beta, 1.0, oracle.flink(cfg["family"], cfg["link"]))
    assert np.max(np.abs(H[:25, :25] - nr["xtwx"])) <= 1e-4 * np.max(np.abs(nr["xtwx"]))
    # and directly: one batch of 40 000 parameter vectors (> 32 768 / P per chunk, > the 16 384-value result buffer)
    mdl = g.Model(gctx, Xw, cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    mdl.set_u(cfg["U"])
    B = np.asfortranarray(beta[:, None] + 1e-3 * rng.standard_normal((25, 40000)))
    ll = mdl.log_likelihood_batch(B, np.ones(40000))
    for k in (0, 17, 16384, 39999):
        assert abs(ll[k] - mdl.log_likelihood(B[:, k], 1.0)) <= 1e-12 * abs(ll[k])
    mdl.close()


def test_entry_points_reuse_their_device_objects_between_calls(gctx, oracle):
    """gmb_set_object_cache: the entry points keep the device objects of the last models / covariance specifications and reuse them on a
    byte-for-byte match of X, Z, y (and cov, data, eff_range) — same results with the cache on and off, in any interleaving of two models, after
    a change of a single response, and after the context that owned them is gone."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    c1 = synth.config2(m=300, seed=31, ncl=8, nt=4, nind=6)
    c2 = synth.config1(m=200)
    out = {}
    for on in (True, False, True):
        g.set_object_cache(on)
        res = []
        for rep in range(2):
            for cfg in (c1, c2):
                start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
                L = synth.dense_chol_D(cfg["cov"], cfg["data"], cfg["theta"])
                u = g.mcmc_sample(cfg["Z"], L, cfg["X"], cfg["y"], cfg["beta"], cfg["family"], cfg["link"], 30, 99, 1.0, 1.0, 0, 500, 20, 0.9, n_chains=4, seed=7)
                fit = g.mcml_optim(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], u, cfg["family"], cfg["link"], start, 0, True)
                ll = g.mvn_ll(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], u)
                res.append((u.copy(), fit["beta"].copy(), fit["theta"].copy(), ll))
        y2 = c1["y"].copy(); y2[3] = 1.0 - y2[3]                     # one response flipped: a different model, not the cached one
        start = np.concatenate([c1["beta"], c1["theta"], [1.0]])
        fit2 = g.mcml_optim(c1["cov"], c1["data"], c1["eff_range"], c1["Z"], c1["X"], y2, res[0][0], c1["family"], c1["link"], start, 0, True)
        out[len(out)] = (res, fit2["beta"].copy())
    g.set_object_cache(True)
    base, b2 = out[0]
    for k in (1, 2):
        res, bb = out[k]
        for a, b in zip(base, res):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
        assert np.array_equal(b2, bb)
    assert not np.array_equal(b2, base[0][1])                       # the flipped response changed the fit
    # the two repetitions inside one setting agree as well (second one runs on the kept objects)
    assert np.array_equal(base[0][1], base[2][1]) and np.array_equal(base[1][1], base[3][1])
    # objects kept for a context go with it
    ctx2 = g.Context(0); ctx2.make_default()
    start = np.concatenate([c2["beta"], c2["theta"], [1.0]])
    f_a = g.mcml_optim(c2["cov"], c2["data"], c2["eff_range"], c2["Z"], c2["X"], c2["y"], base[1][0], c2["family"], c2["link"], start, 0, True)
    gctx.make_default()
    ctx2.close()
    f_b = g.mcml_optim(c2["cov"], c2["data"], c2["eff_range"], c2["Z"], c2["X"], c2["y"], base[1][0], c2["family"], c2["link"], start, 0, True)
    assert np.array_equal(f_a["beta"], f_b["beta"]) and np.array_equal(f_a["theta"], f_b["theta"])
