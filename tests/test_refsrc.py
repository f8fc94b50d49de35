"""The reference's OWN entry-point bodies — src/mcml_optim.cpp, src/mcml_full.cpp, src/mcml_la.cpp compiled unmodified against oracle/shim
(oracle/_ref/librefsrc.so, oracle/refsrc_driver.cpp) — against the oracle the GPU tests use as their checker.  With these, the chain
"CUDA library == oracle (GPU tests) and oracle == the reference's own source (here, on the CPU)" covers the entry points, the MCML loop and
the Laplace fits, not only the numeric headers.  Stand-ins below the reference's code: Eigen / Rcpp / glmmrBase / rminqa (see the driver)."""
import os

import numpy as np
import pytest
from scipy.optimize import minimize

from glmmrmcml_b200 import synth

refsrc = pytest.importorskip("oracle.refsrc")
pytestmark = pytest.mark.skipif(not refsrc.available(), reason="oracle/_ref/librefsrc.so not built (needs /root/reference)")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_rct(m=60, seed=11):
    return synth.config2(m=m, seed=seed, ncl=8, nt=4, nind=6)      # n = 192, P = 5, Q = 32, 8 blocks of 4 (as tests/test_gpu_entry_points.py)


def _args(cfg):
    return (cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["U"], cfg["family"], cfg["link"])


def test_stateless_exports_are_the_oracle_formulas(oracle):
    """mvn_ll (src/mcml_optim.cpp:406-414), aic_mcml (:356-392) and mcmc_sample (src/mcml_full.cpp:314-338): bit for bit."""
    for cfg in (small_rct(), synth.config3(nloc=31, m=9, seed=7), synth.config4(ncl=9, nt=4, k=3, m=21, seed=8)):
        cov = (cfg["cov"], cfg["data"], cfg["eff_range"])
        fl = oracle.flink(cfg["family"], cfg["link"])
        X, Z, y, U = cfg["X"], cfg["Z"], cfg["y"], cfg["U"]
        d = oracle.mvn_loglik(*cov, cfg["theta"], U, faithful=True)
        assert refsrc.mvn_ll(*cov, cfg["theta"], U) == d
        gauss = cfg["family"] == "gaussian"
        bp = np.concatenate([cfg["beta"], [0.7]]) if gauss else cfg["beta"]
        ll = oracle.loglik_faithful(X, Z, U, y, cfg["beta"], 0.7 if gauss else 0.0, fl)
        want = -2 * (ll + d) + 2 * (bp.size + cfg["theta"].size)                       # :391
        got = refsrc.aic_mcml(*_args(cfg), bp, cfg["theta"])
        assert abs(got - want) <= 1e-13 * abs(want)
        u = refsrc.mcmc_sample(Z, cfg["L"], X, y, cfg["beta"], cfg["family"], cfg["link"], 20, 10, 0.5, 1.0, 0, 500, 15, 0.9, seed=4242)
        ch = oracle.hmc_chain(oracle.gemm(Z, cfg["L"]), cfg["L"], X @ cfg["beta"], y, 1.0, fl, 20, 10, 0.5, 15, 0.9, 4242, chain=0)
        assert u.shape == (cfg["Q"], 11) and np.max(np.abs(u - ch["u"])) <= 1e-12 * max(1.0, np.max(np.abs(u)))


def _objectives(cfg, oracle):
    fl = oracle.flink(cfg["family"], cfg["link"])
    zd = oracle.gemm(cfg["Z"], cfg["U"])
    L_obj = lambda beta, sigma=1.0: -oracle.loglik_zd(zd, cfg["X"] @ beta, cfg["y"], sigma, fl)       # L_likelihood, likelihood.h:57-64

    def D_obj(theta):                                                                                 # D_likelihood, likelihood.h:40-45
        if np.any(np.asarray(theta) < 1e-6):
            return 1e300
        v = -oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], np.asarray(theta, dtype=np.float64), cfg["U"])
        return v if np.isfinite(v) else 1e300
    return L_obj, D_obj


def test_mcml_optim_of_the_reference_against_the_oracle_objectives(oracle):
    """mcml_optim (src/mcml_optim.cpp:35-68): the MCNR beta is mcmloptim::mcnr's (oracle: 1e-10), the MCEM beta and theta are the minimisers
    of the oracle's L / D objectives (scipy; the stand-in optimiser of oracle/shim/rbobyqa.h is compared at optimiser tolerance)."""
    cfg = small_rct()
    P = cfg["P"]
    L_obj, D_obj = _objectives(cfg, oracle)
    start = np.concatenate([cfg["beta"] * 0.8, cfg["theta"] * 1.2, [1.0]])
    fl = oracle.flink(cfg["family"], cfg["link"])
    nr = refsrc.mcml_optim(*_args(cfg), start, 0, True)
    o = oracle.mcnr(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], start[:P], 1.0, fl)
    assert np.max(np.abs(nr["beta"] - (start[:P] + o["beta_incr"]))) <= 1e-10 and nr["sigma"] == o["sigma"]
    th = minimize(D_obj, start[P:P + 2], method="Nelder-Mead", options=dict(xatol=1e-10, fatol=1e-14, maxiter=4000)).x
    assert np.max(np.abs(nr["theta"] - th)) <= 1e-5, (nr["theta"], th)
    em = refsrc.mcml_optim(*_args(cfg), start, 0, False)
    b = minimize(L_obj, start[:P], method="BFGS", options=dict(gtol=1e-9)).x
    assert np.max(np.abs(em["beta"] - b)) <= 1e-5 and np.max(np.abs(em["theta"] - th)) <= 1e-5 and em["sigma"] == 0.0   # sigma_ = 0, mcmloptim.h:30


def test_mcml_hess_of_the_reference_is_the_library_stencil_on_the_oracle_objective(oracle):
    """mcml_hess (src/mcml_optim.cpp:263-285 -> f_hess, mcmloptim.h:333-355 -> optimhess) against the product's host-side stencil
    gmb_fd_hessian driven by the oracle's F objective: same 4 k^2 points, same symmetrisation."""
    import glmmrmcml_b200 as g
    cfg = small_rct()
    P = cfg["P"]
    L_obj, D_obj = _objectives(cfg, oracle)
    F = lambda Xp: np.array([L_obj(Xp[:P, k]) + D_obj(Xp[P:, k]) for k in range(Xp.shape[1])])
    x0 = np.concatenate([cfg["beta"], cfg["theta"]])
    tol = 1e-3
    H = refsrc.mcml_hess(*_args(cfg), x0, tol, 0)
    lower = np.concatenate([np.full(P, -np.inf), np.full(2, 1e-6)])
    Hw, nfev = g.fd_hessian(F, x0, tol, lower=lower, upper=np.full(P + 2, np.inf), usebounds=True)
    assert H.shape == (P + 2, P + 2) and np.array_equal(H, H.T)
    assert np.max(np.abs(H - Hw)) <= 1e-7 * np.max(np.abs(Hw)), np.max(np.abs(H - Hw)) / np.max(np.abs(Hw))


def test_mcml_simlik_of_the_reference_against_the_oracle_objective(oracle):
    """mcml_simlik (src/mcml_optim.cpp:90-117): joint minimisation of F_likelihood with importance weights (likelihood.h:88-108) — on a model
    small enough for exp(ll + logl) to stay in range the optimum is that of L + D (the denominator is a constant)."""
    cfg = small_rct(m=40)
    P = cfg["P"]
    L_obj, D_obj = _objectives(cfg, oracle)
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    fit = refsrc.mcml_simlik(*_args(cfg), start, 0)
    r = minimize(lambda p: L_obj(p[:P]) + D_obj(p[P:]), start[:P + 2], method="BFGS", options=dict(gtol=1e-8))
    assert np.max(np.abs(np.concatenate([fit["beta"], fit["theta"]]) - r.x)) <= 5e-5


FULL_CASES = {
    "binomial_mcnr": (lambda: small_rct(m=8), True), "binomial_mcem": (lambda: small_rct(m=8), False),
    "gaussian_mcnr": (lambda: synth.config3(nloc=25, m=8, seed=7), True), "gaussian_mcem": (lambda: synth.config3(nloc=25, m=8, seed=7), False),
    "poisson_mcnr": (lambda: synth.config4(ncl=9, nt=4, k=3, m=8, seed=8), True),
}


@pytest.mark.parametrize("name", list(FULL_CASES))
def test_the_reference_mcml_full_loop_is_the_oracle_loop(name, oracle):
    """src/mcml_full.cpp:41-148 as written (sampler object, niter_ = m quirk, abs() convergence test, refresh of L / xb / var_par) against
    oracle/mcml_loop.py — the checker of tests/test_gpu_fit_parity.py — on the same Philox stream, several iterations, two seeds."""
    from oracle import mcml_loop
    make, mcnr = FULL_CASES[name]
    cfg = make()
    gauss = cfg["family"] == "gaussian"
    start = np.concatenate([cfg["beta"] * 0.8, cfg["theta"] * 1.2, [0.8 if gauss else 1.0]])
    a = (cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start)
    kw = dict(mcnr=mcnr, m=40, maxiter=4, warmup=40, tol=1e-3, lam=0.5, maxsteps=15, target_accept=0.9)
    for seed in (5, 7):
        f = refsrc.mcml_full(*a, seed=seed, **kw)
        o = mcml_loop.mcml_full(*a, seed=seed, **kw)
        assert f["converged"] == o["converged"]
        assert np.max(np.abs(f["beta"] - o["beta"])) <= 1e-6, (seed, f["beta"], o["beta"])
        assert np.max(np.abs(f["theta"] - o["theta"])) <= 1e-6, (seed, f["theta"], o["theta"])
        assert abs(f["sigma"] - o["sigma"]) <= 1e-6
        assert f["u"].shape == o["u"].shape == (cfg["Q"], 41) and np.max(np.abs(f["u"] - o["u"])) <= 1e-6


def test_the_reference_loop_at_the_gpu_fit_parity_settings(oracle):
    """One seed of tests/test_gpu_fit_parity.py::test_c2_mcnr_fit_follows_the_oracle_loop_over_10_seeds, configuration and settings
    unchanged: the oracle loop the GPU test compares the library with IS the reference's own loop (all ten seeds of both configurations:
    tests/golden/REFSRC_mcml_full.npz, made by tests/golden/make_golden_refsrc.py)."""
    from oracle import mcml_loop
    gold = np.load(os.path.join(ROOT, "tests", "golden", "REFSRC_mcml_full.npz"))
    cfg = synth.config2(m=8)
    start = np.concatenate([cfg["beta"] * 0.8, [0.3, 0.6], [1.0]])
    assert np.array_equal(start, gold["C2_mcnr_start"])
    kw = dict(mcnr=True, m=250, maxiter=6, warmup=150, tol=1e-2, lam=5.0, maxsteps=100, target_accept=0.95)
    a = (cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start)
    k = 3
    seed = int(gold["seeds"][k])
    o = mcml_loop.mcml_full(*a, seed=seed, **kw)
    assert np.max(np.abs(o["beta"] - gold["C2_mcnr_beta"][k])) <= 1e-6 and np.max(np.abs(o["theta"] - gold["C2_mcnr_theta"][k])) <= 1e-6
    assert bool(gold["C2_mcnr_converged"][k]) == o["converged"]
    assert np.max(np.abs(o["u"][:, -1] - gold["C2_mcnr_u_last_column"][k])) <= 1e-6


def test_laplace_fits_of_the_reference_against_the_staged_replay(oracle):
    """mcml_la_nr / mcml_la with maxiter = 1 (src/mcml_la.cpp:178-290, :28-155): the staged replay on the numpy oracle that the GPU tests
    (tests/test_laplace.py) hold the CUDA library to, held here to the reference's own function bodies."""
    from oracle import laplace
    # --- mcml_la_nr, binomial ---
    cfg = small_rct(m=4)
    fam, link = "binomial", "logit"
    fl = oracle.flink(fam, link)
    P, Q = cfg["P"], cfg["Q"]
    X, Z, y = cfg["X"], cfg["Z"], cfg["y"]
    cov, data, eff = cfg["cov"], cfg["data"], cfg["eff_range"]
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    fit = refsrc.mcml_la_nr(cov, data, eff, Z, X, y, fam, link, start, usehess=True, tol=1e-4, maxiter=1)
    L0 = oracle.genD(cov, data, eff, cfg["theta"], chol=True)
    v0 = np.zeros(Q)
    W0 = laplace.w_diag(X @ cfg["beta"], (Z @ L0) @ v0, 1.0, fam, link)
    b1, v1, _ = laplace.mcnr_b(X, Z, L0, L0 @ L0.T, y, cfg["beta"], v0, W0, 1.0, fam, link, fl)
    W1 = laplace.w_diag(X @ b1, (Z @ L0) @ v1, 1.0, fam, link)
    guard = lambda th: np.all(np.asarray(th) > 1e-6) and th[1] < 0.999
    t1 = minimize(lambda th: laplace.la_likelihood_cov(th, cov, data, eff, Z, X @ b1, y, v1, W1, fam, fl, 1.0) if guard(th) else 1e300,
                  cfg["theta"], method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=4000)).x
    rC = minimize(lambda p: laplace.la_likelihood_btheta(p, cov, data, eff, Z, X, y, v1, fam, link, fl, 1.0) if guard(p[P:]) else 1e300,
                  np.concatenate([b1, t1]), method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=20000, maxfev=40000))
    u_want = oracle.genD(cov, data, eff, t1, chol=True) @ v1
    assert np.max(np.abs(fit["u"].ravel() - u_want)) <= 1e-3 * max(1.0, np.max(np.abs(u_want)))
    got = np.concatenate([fit["beta"], fit["theta"]])
    f_got = laplace.la_likelihood_btheta(got, cov, data, eff, Z, X, y, v1, fam, link, fl, 1.0)
    assert f_got <= rC.fun + 1e-6 * abs(rC.fun) and np.max(np.abs(got - rC.x)) <= 2e-2
    assert np.all(fit["se"][: P + 2] > 0) and fit["se"][-1] == 0.0
    # --- mcml_la, poisson ---
    cfg = synth.config4(ncl=6, nt=4, k=4, m=4)
    fam, link = "poisson", "log"
    fl = oracle.flink(fam, link)
    P, Q = cfg["P"], cfg["Q"]
    X, Z, y = cfg["X"], cfg["Z"], cfg["y"]
    cov, data, eff = cfg["cov"], cfg["data"], cfg["eff_range"]
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    fit = refsrc.mcml_la(cov, data, eff, Z, X, y, fam, link, start, usehess=False, tol=1e-3, maxiter=1)
    L0 = oracle.genD(cov, data, eff, cfg["theta"], chol=True)
    rA = minimize(lambda p: laplace.la_likelihood(p, X, Z @ L0, y, 1.0, fl), np.concatenate([cfg["beta"], np.zeros(Q)]), method="BFGS", options=dict(gtol=1e-8))
    b1, v1 = rA.x[:P], rA.x[P:]
    W1 = laplace.w_diag(X @ b1, Z @ v1, 1.0, fam, link)
    t1 = minimize(lambda th: laplace.la_likelihood_cov(th, cov, data, eff, Z, X @ b1, y, v1, W1, fam, fl, 1.0) if guard(th) else 1e300,
                  cfg["theta"], method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=4000)).x
    rC = minimize(lambda p: laplace.la_likelihood_btheta(p, cov, data, eff, Z, X, y, v1, fam, link, fl, 1.0) if guard(p[P:]) else 1e300,
                  np.concatenate([b1, t1]), method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=20000, maxfev=40000))
    u_want = oracle.genD(cov, data, eff, t1, chol=True) @ v1
    assert np.max(np.abs(fit["u"].ravel() - u_want)) <= 1e-3 * max(1.0, np.max(np.abs(u_want)))
    got = np.concatenate([fit["beta"], fit["theta"]])
    f_got = laplace.la_likelihood_btheta(got, cov, data, eff, Z, X, y, v1, fam, link, fl, 1.0)
    assert f_got <= rC.fun + 1e-6 * abs(rC.fun) and np.max(np.abs(got - rC.x)) <= 2e-2


def test_golden_entry_point_values_are_the_reference_entry_points():
    """tests/test_golden.py (GPU) asserts gmb_mvn_ll, gmb_aic_mcml and gmb_mcml_optim(mcnr) against numbers derived from the golden files
    (outputs of the reference's HEADERS).  Here the reference's own ENTRY POINTS (src/mcml_optim.cpp) are run on the golden inputs and give
    those same numbers — so that GPU test is a comparison with mvn_ll(), aic_mcml() and mcml_optim() of the reference themselves."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "C*_*.npz")))
    assert len(files) == 5
    for path in files:
        g = np.load(path)
        fam, link = str(g["family"]), str(g["link"])
        X, Z, y, U = g["X"], g["Z"], g["y"], g["U"]
        cov = (g["cov"], g["data"], g["eff_range"])
        assert refsrc.mvn_ll(*cov, g["theta"], U) == g["mvn_ll"][0]
        bp = np.concatenate([g["beta"], [1.0]]) if fam == "gaussian" else g["beta"]
        want = -2 * (-(g["objectives"][0]) - g["objectives"][1]) + 2 * (bp.size + g["theta"].size)          # as in tests/test_golden.py
        got = refsrc.aic_mcml(*cov, Z, X, y, U, fam, link, bp, g["theta"])
        # (aic_mcml evaluates non-gaussian families at var_par = 0, src/mcml_optim.cpp:380-384, the golden objectives at 1: the same value)
        assert abs(got - want) <= 1e-12 * abs(want)
        opt = refsrc.mcml_optim(*cov, Z, X, y, U, fam, link, np.concatenate([g["beta"], g["theta"], [1.0]]), 0, True)
        assert np.max(np.abs(opt["beta"] - g["mcnr_beta"])) <= 1e-12 * max(1.0, np.max(np.abs(g["mcnr_beta"])))
        assert opt["sigma"] == float(g["mcnr_sigma"])


def test_the_product_optimiser_on_the_oracle_objectives_reaches_the_reference_m_step(oracle):
    """gmb_minimize_bounded (csrc/optim.cpp — host code, callable without a GPU; the library's l_optim / d_optim call it with exactly these
    settings, csrc/api.cpp:168,175) driven by the ORACLE's objectives reproduces the M-step of the reference's own mcml_optim (its l_optim /
    d_optim with the stand-in optimiser): both optimisers stop at the same optimum, far below the MCML tolerance."""
    import glmmrmcml_b200 as g
    for cfg in (small_rct(), synth.config4(ncl=9, nt=4, k=3, m=21, seed=8)):
        P = cfg["P"]
        L_obj, D_obj = _objectives(cfg, oracle)
        start = np.concatenate([cfg["beta"] * 0.8, cfg["theta"] * 1.2, [1.0]])
        ref_fit = refsrc.mcml_optim(*_args(cfg), start, 0, False)
        b = g.minimize_bounded(lambda Xp: np.array([L_obj(Xp[:, k]) for k in range(Xp.shape[1])]), start[:P], xtol=1e-8, maxit=200)
        t = g.minimize_bounded(lambda Xp: np.array([D_obj(Xp[:, k]) for k in range(Xp.shape[1])]), start[P:P + 2], lower=np.full(2, 1e-6), xtol=1e-8, maxit=200)
        assert np.max(np.abs(b["x"] - ref_fit["beta"])) <= 2e-5, (b["x"], ref_fit["beta"])
        assert np.max(np.abs(t["x"] - ref_fit["theta"])) <= 2e-5, (t["x"], ref_fit["theta"])
