"""Laplace-approximation path (mcml_la / mcml_la_nr, src/mcml_la.cpp; likelihood.h:112-230; mcmloptim.h:116-195, 238-293).

CPU: the numpy restatement oracle/laplace.py reproduces the golden vectors written by the reference's own functors
(tests/golden/make_golden_laplace.py) and, where oracle/_ref is present, libref itself.
GPU: the device objectives and the Newton step reproduce the golden vectors to 1e-10; the two fits are checked against
scipy optimisers of the oracle's objectives."""
import glob
import os

import numpy as np
import pytest

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "LA_*.npz")))
FLINK = {("poisson", "log"): 1, ("binomial", "logit"): 3, ("gaussian", "identity"): 7}


def load(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    g["family"], g["link"], g["sigma"] = str(g["family"]), str(g["link"]), float(g["sigma"])
    return g


def oracle_state(g, oracle, use_l):
    from oracle import laplace
    fam, link = g["family"], g["link"]
    fl = FLINK[(fam, link)]
    X, Z, y, beta, theta, v = g["X"], g["Z"], g["y"], g["beta"], g["theta"], g["v"]
    L = oracle.genD(g["cov"], g["data"], g["eff_range"], theta, chol=True)
    vp = g["sigma"] if fam == "gaussian" else 1.0
    xb = X @ beta
    W = laplace.w_diag(xb, ((Z @ L) if use_l else Z) @ v, vp, fam, link)
    pt = np.concatenate([theta, [g["sigma"]]]) if fam == "gaussian" else theta
    pbt = np.concatenate([beta, pt])
    obj = np.array([laplace.la_likelihood(np.concatenate([beta, v]), X, Z @ L, y, vp, fl),
                    laplace.la_likelihood_cov(pt, g["cov"], g["data"], g["eff_range"], Z, xb, y, v, W, fam, fl, vp),
                    laplace.la_likelihood_btheta(pbt, g["cov"], g["data"], g["eff_range"], Z, X, y, v, fam, link, fl, vp)])
    nr = laplace.mcnr_b(X, Z, L, L @ L.T, y, beta, v, W, vp, fam, link, fl)
    return obj, nr


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_numpy_oracle_reproduces_reference_laplace_outputs(path, oracle):
    g = load(path)
    for use_l in (0, 1):
        obj, (bn, vn, sn) = oracle_state(g, oracle, bool(use_l))
        assert np.max(np.abs(obj - g[f"obj_{use_l}"]) / np.abs(g[f"obj_{use_l}"])) <= 1e-13
        assert np.max(np.abs(bn - g[f"beta_nr_{use_l}"])) <= 1e-12 * max(1.0, np.max(np.abs(bn)))
        assert np.max(np.abs(vn - g[f"v_nr_{use_l}"])) <= 1e-12 * max(1.0, np.max(np.abs(vn)))
        assert abs(sn - float(g[f"sigma_nr_{use_l}"])) <= 1e-13


def test_golden_laplace_vectors_are_libref_outputs(oracle):
    """Where the reference tree was available at build time: the golden files are what libref computes now."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (no reference tree)")
    for path in GOLD:
        g = load(path)
        r = ref.la_objectives(g["cov"], g["data"], g["eff_range"], g["X"], g["Z"], g["y"], g["family"], g["link"], g["beta"], g["theta"], g["v"], g["sigma"], False)
        assert np.array_equal(np.array([r["la"], r["la_cov"], r["la_btheta"]]), g["obj_0"])
        assert np.array_equal(r["v_nr"], g["v_nr_0"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_reproduces_reference_laplace_outputs(path, gctx):
    import glmmrmcml_b200 as g_
    gctx.make_default()
    g = load(path)
    for use_l in (0, 1):
        r = g_.la_objectives(g["cov"], g["data"], g["eff_range"], g["Z"], g["X"], g["y"], g["family"], g["link"], g["beta"], g["theta"], g["v"],
                             g["sigma"], bool(use_l))
        got = np.array([r["la"], r["la_cov"], r["la_btheta"]])
        want = g[f"obj_{use_l}"]
        assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-10, (got, want)
        assert np.max(np.abs(r["beta_nr"] - g[f"beta_nr_{use_l}"])) <= 1e-9 * max(1.0, np.max(np.abs(g[f"beta_nr_{use_l}"])))
        assert np.max(np.abs(r["v_nr"] - g[f"v_nr_{use_l}"])) <= 1e-9 * max(1.0, np.max(np.abs(g[f"v_nr_{use_l}"])))
        assert abs(r["sigma_nr"] - float(g[f"sigma_nr_{use_l}"])) <= 1e-10


@pytest.mark.gpu
def test_mcml_la_nr_one_iteration_matches_staged_replay_on_the_oracle(gctx, oracle):
    """mcml_la_nr with maxiter = 1 (src/mcml_la.cpp:199,216-219,260): W at xb + Z L v, one mcnr_b step from v = 0, theta on
    LA_likelihood_cov, then (beta, theta) on LA_likelihood_btheta; u = L(theta1) v1.  Replayed with the numpy oracle + scipy."""
    import glmmrmcml_b200 as g_
    from oracle import laplace
    from scipy.optimize import minimize
    from glmmrmcml_b200 import synth
    gctx.make_default()
    cfg = synth.config2(m=4, seed=11, ncl=8, nt=4, nind=6)
    fam, link = "binomial", "logit"
    fl = FLINK[(fam, link)]
    P, Q = cfg["P"], cfg["Q"]
    X, Z, y = cfg["X"], cfg["Z"], cfg["y"]
    cov, data, eff = cfg["cov"], cfg["data"], cfg["eff_range"]
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    fit = g_.mcml_la_nr(cov, data, eff, Z, X, y, fam, link, start, usehess=True, tol=1e-4, verbose=False, maxiter=1)
    assert fit["u"].shape == (Q, 1) and fit["iter"] == 1
    L0 = oracle.genD(cov, data, eff, cfg["theta"], chol=True)
    v0 = np.zeros(Q)
    W0 = laplace.w_diag(X @ cfg["beta"], (Z @ L0) @ v0, 1.0, fam, link)
    b1, v1, _ = laplace.mcnr_b(X, Z, L0, L0 @ L0.T, y, cfg["beta"], v0, W0, 1.0, fam, link, fl)
    W1 = laplace.w_diag(X @ b1, (Z @ L0) @ v1, 1.0, fam, link)               # update_W(0, true)
    guard = lambda th: np.all(np.asarray(th) > 1e-6) and th[1] < 0.999
    rB = minimize(lambda th: laplace.la_likelihood_cov(th, cov, data, eff, Z, X @ b1, y, v1, W1, fam, fl, 1.0) if guard(th) else 1e300,
                  cfg["theta"], method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=4000))
    t1 = rB.x
    rC = minimize(lambda p: laplace.la_likelihood_btheta(p, cov, data, eff, Z, X, y, v1, fam, link, fl, 1.0) if guard(p[P:]) else 1e300,
                  np.concatenate([b1, t1]), method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=20000, maxfev=40000))
    u_want = oracle.genD(cov, data, eff, t1, chol=True) @ v1
    assert np.max(np.abs(fit["u"].ravel() - u_want)) <= 1e-3 * max(1.0, np.max(np.abs(u_want))), np.max(np.abs(fit["u"].ravel() - u_want))
    got = np.concatenate([fit["beta"], fit["theta"]])
    f_got = laplace.la_likelihood_btheta(got, cov, data, eff, Z, X, y, v1, fam, link, fl, 1.0)
    assert f_got <= rC.fun + 1e-6 * abs(rC.fun), (f_got, rC.fun)
    assert np.max(np.abs(got - rC.x)) <= 2e-2, (got, rC.x)
    # hess_la standard errors (mcmloptim.h:179-195): positive for (beta, theta), zero in the unused last slot
    assert np.all(np.isfinite(fit["se"][: P + 2])) and np.all(fit["se"][: P + 2] > 0) and fit["se"][-1] == 0.0
    # and a full run converges to finite estimates
    full = g_.mcml_la_nr(cov, data, eff, Z, X, y, fam, link, start, usehess=False, tol=1e-4, verbose=False, maxiter=25)
    assert np.all(np.isfinite(full["beta"])) and np.all(full["theta"] > 0) and full["iter"] >= 2


@pytest.mark.gpu
def test_mcml_la_one_iteration_matches_staged_scipy_on_oracle_objectives(gctx, oracle):
    """mcml_la with maxiter = 1 is three optimisations in sequence (src/mcml_la.cpp:64-68,107): (beta, v) on LA_likelihood at theta0,
    theta on LA_likelihood_cov, (beta, theta) on LA_likelihood_btheta; u = L(theta1) v1.  Each stage is replayed with scipy on the
    oracle's objectives."""
    import glmmrmcml_b200 as g_
    from oracle import laplace
    from scipy.optimize import minimize
    from glmmrmcml_b200 import synth
    gctx.make_default()
    cfg = synth.config4(ncl=6, nt=4, k=4, m=4)
    fam, link = "poisson", "log"
    fl = FLINK[(fam, link)]
    P, Q = cfg["P"], cfg["Q"]
    X, Z, y = cfg["X"], cfg["Z"], cfg["y"]
    cov, data, eff = cfg["cov"], cfg["data"], cfg["eff_range"]
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    fit = g_.mcml_la(cov, data, eff, Z, X, y, fam, link, start, usehess=False, tol=1e-3, verbose=False, maxiter=1)
    assert np.all(np.isfinite(fit["beta"])) and np.all(fit["theta"] > 0) and fit["iter"] == 1
    L0 = oracle.genD(cov, data, eff, cfg["theta"], chol=True)
    rA = minimize(lambda p: laplace.la_likelihood(p, X, Z @ L0, y, 1.0, fl), np.concatenate([cfg["beta"], np.zeros(Q)]), method="BFGS",
                  options=dict(gtol=1e-8))
    b1, v1 = rA.x[:P], rA.x[P:]
    W1 = laplace.w_diag(X @ b1, Z @ v1, 1.0, fam, link)                       # update_W(): Z v
    guard = lambda th: np.all(np.asarray(th) > 1e-6) and th[1] < 0.999
    rB = minimize(lambda th: laplace.la_likelihood_cov(th, cov, data, eff, Z, X @ b1, y, v1, W1, fam, fl, 1.0) if guard(th) else 1e300,
                  cfg["theta"], method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=4000))
    t1 = rB.x
    rC = minimize(lambda p: laplace.la_likelihood_btheta(p, cov, data, eff, Z, X, y, v1, fam, link, fl, 1.0) if guard(p[P:]) else 1e300,
                  np.concatenate([b1, t1]), method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-13, maxiter=20000, maxfev=40000))
    u_want = oracle.genD(cov, data, eff, t1, chol=True) @ v1
    assert np.max(np.abs(fit["u"].ravel() - u_want)) <= 1e-3 * max(1.0, np.max(np.abs(u_want))), np.max(np.abs(fit["u"].ravel() - u_want))
    got = np.concatenate([fit["beta"], fit["theta"]])
    f_got = laplace.la_likelihood_btheta(got, cov, data, eff, Z, X, y, v1, fam, link, fl, 1.0)
    assert f_got <= rC.fun + 1e-6 * abs(rC.fun), (f_got, rC.fun)             # at least as good as scipy's optimum of the last stage
    assert np.max(np.abs(got - rC.x)) <= 2e-2, (got, rC.x)
