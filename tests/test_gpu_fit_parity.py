"""Final-estimate parity (BASELINE.json north_star: "the final MCML estimates must agree with the reference to its own tol") and
sampler posterior moments against the oracle chain.

gmb_mcml_full with n_chains = 1 is the reference's loop with its single chain (src/mcml_full.cpp:83-126, mhmcmc.h:121-157); the
oracle-side loop (oracle/mcml_loop.py: oracle chain + scipy bounded minimisers on the oracle's objectives) runs on the SAME Philox
stream, so the two fits can be compared iterate for iterate and seed for seed — not only in the mean over seeds: over >= 10 seeds per
configuration every seed's estimates agree far below the reference's tolerance (C1: tol 5e-3, MCEM; C2: tol 1e-2, MCNR)."""
import os

import numpy as np
import pytest

from glmmrmcml_b200 import synth

pytestmark = pytest.mark.gpu
SEEDS = list(range(101, 111))          # 10 seeds


def _refsrc_gold(name, seed, cfg=None):
    """The fit of the reference's OWN src/mcml_full.cpp (compiled unmodified against oracle/shim: oracle/refsrc_driver.cpp) for this
    configuration, these settings and this seed's Philox stream — tests/golden/REFSRC_mcml_full.npz, made by
    tests/golden/make_golden_refsrc.py.  On the CPU, tests/test_refsrc.py and tools/check_refsrc_golden.py hold the oracle loop to the
    same file at 1e-6, so the tolerances below are the oracle-loop tolerances plus a margin."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "REFSRC_mcml_full.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    if cfg is not None and not (np.array_equal(z[name + "_y"], cfg["y"]) and np.array_equal(z[name + "_X"], cfg["X"])):
        return None                     # the synthetic-data generator no longer produces the data set the golden fits were made on
    k = [int(s) for s in z["seeds"]].index(int(seed))
    return dict(beta=z[name + "_beta"][k], theta=z[name + "_theta"][k], converged=bool(z[name + "_converged"][k]), u_last=z[name + "_u_last_column"][k],
                start=z[name + "_start"])


def _fit_pair(g, mcml_loop, cfg, start, seed, **kw):
    a = (cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start)
    dev = g.mcml_full(*a, verbose=False, n_chains=1, seed=seed, **kw)
    orc = mcml_loop.mcml_full(*a, seed=seed, **kw)
    return dev, orc


def test_c2_mcnr_fit_follows_the_oracle_loop_over_10_seeds(gctx, oracle):
    """C2 (cluster RCT, gr(cl)*ar1(t), binomial-logit, MCNR) at the default tol = 1e-2."""
    import glmmrmcml_b200 as g
    from oracle import mcml_loop
    gctx.make_default()
    cfg = synth.config2(m=8)
    start = np.concatenate([cfg["beta"] * 0.8, [0.3, 0.6], [1.0]])
    kw = dict(mcnr=True, m=250, maxiter=6, warmup=150, tol=1e-2, lam=5.0, maxsteps=100, target_accept=0.95)
    db, dt, ob, ot = [], [], [], []
    for seed in SEEDS:
        dev, orc = _fit_pair(g, mcml_loop, cfg, start, seed, **kw)
        assert dev["iter"] == orc["iter"] and dev["converged"] == orc["converged"], (seed, dev["iter"], orc["iter"])
        assert np.max(np.abs(dev["beta"] - orc["beta"])) <= 1e-4, (seed, dev["beta"], orc["beta"])
        assert np.max(np.abs(dev["theta"] - orc["theta"])) <= 1e-4, (seed, dev["theta"], orc["theta"])
        assert dev["u"].shape == orc["u"].shape == (cfg["Q"], 251)
        assert np.max(np.abs(dev["u"] - orc["u"])) <= 1e-4          # the last iteration's samples: the same chain
        ref = _refsrc_gold("C2_mcnr", seed, cfg)                         # and against the reference's own src/mcml_full.cpp, directly
        if ref is not None:
            assert np.array_equal(ref["start"], start) and dev["converged"] == ref["converged"]
            assert np.max(np.abs(dev["beta"] - ref["beta"])) <= 2e-4 and np.max(np.abs(dev["theta"] - ref["theta"])) <= 2e-4, (seed, dev["beta"], ref["beta"])
            assert np.max(np.abs(dev["u"][:, -1] - ref["u_last"])) <= 2e-4
        db.append(dev["beta"]); dt.append(dev["theta"]); ob.append(orc["beta"]); ot.append(orc["theta"])
    # SURVEY §7's statement of the same thing: means over the seeds agree within the reference's tolerance
    assert np.max(np.abs(np.mean(db, 0) - np.mean(ob, 0))) <= 1e-2 and np.max(np.abs(np.mean(dt, 0) - np.mean(ot, 0))) <= 1e-2
    # Monte-Carlo spread of the iterates across seeds, for the record (the reference's stopping rule compares successive iterates)
    print("C2 sd over seeds: beta", np.std(db, 0).round(4), "theta", np.std(dt, 0).round(4))


def test_c1_mcem_fit_follows_the_oracle_loop_over_10_seeds(gctx, oracle):
    """C1 (README cluster RCT, (1|gr(cl)) + (1|gr(cl,t)), binomial-logit, MCEM m = 250) at the README's tol = 5e-3."""
    import glmmrmcml_b200 as g
    from oracle import mcml_loop
    gctx.make_default()
    cfg = synth.config1(m=8)
    start = np.concatenate([cfg["beta"] * 0.8, [0.3, 0.2], [1.0]])
    kw = dict(mcnr=False, m=250, maxiter=4, warmup=150, tol=5e-3, lam=5.0, maxsteps=100, target_accept=0.95)
    db, ob = [], []
    for seed in SEEDS:
        dev, orc = _fit_pair(g, mcml_loop, cfg, start, seed, **kw)
        assert dev["iter"] == orc["iter"] and dev["converged"] == orc["converged"]
        assert np.max(np.abs(dev["beta"] - orc["beta"])) <= 5e-4, (seed, dev["beta"], orc["beta"])
        assert np.max(np.abs(dev["theta"] - orc["theta"])) <= 5e-4, (seed, dev["theta"], orc["theta"])
        ref = _refsrc_gold("C1_mcem", seed, cfg)                         # the reference's own src/mcml_full.cpp (MCEM: optimiser tolerance)
        if ref is not None:
            assert np.array_equal(ref["start"], start) and dev["converged"] == ref["converged"]
            assert np.max(np.abs(dev["beta"] - ref["beta"])) <= 6e-4 and np.max(np.abs(dev["theta"] - ref["theta"])) <= 6e-4, (seed, dev["beta"], ref["beta"])
        db.append(np.concatenate([dev["beta"], dev["theta"]])); ob.append(np.concatenate([orc["beta"], orc["theta"]]))
    assert np.max(np.abs(np.mean(db, 0) - np.mean(ob, 0))) <= 5e-3
    # with m = 250 the Monte-Carlo noise of an iterate is larger than tol = 5e-3 — for the oracle loop exactly as for the device loop
    # (why neither meets the README's tolerance within 30 iterations on this data set)
    sd = np.std(ob, 0)
    print("C1 sd of the iterate over seeds (oracle loop):", sd.round(4))
    assert np.max(sd) > 5e-3


@pytest.mark.parametrize("family", ["binomial", "poisson"])
def test_posterior_moments_match_long_oracle_chains(gctx, oracle, family):
    """Sampler parity in distribution (north_star: posterior moments of u within Monte-Carlo error): mean and variance of every
    random effect from the batched device sampler against 10 independent replications of the oracle's sequential chain."""
    import glmmrmcml_b200 as g
    cfg = synth.config2(m=8, seed=21, ncl=8, nt=4, nind=6) if family == "binomial" else synth.config4(ncl=8, nt=4, k=3, m=8, seed=23)
    fl = oracle.flink(cfg["family"], cfg["link"])
    ZL = cfg["Z"] @ cfg["L"]; xb = cfg["X"] @ cfg["beta"]; Q = cfg["Q"]
    hm = dict(lam=2.0, max_steps=40, target_accept=0.9)
    reps = []
    for r in range(10):
        ch = oracle.hmc_chain(ZL, cfg["L"], xb, cfg["y"], 1.0, fl, 300, 1500, hm["lam"], hm["max_steps"], hm["target_accept"], 900 + r, chain=0)
        reps.append(ch["u"][:, 1:])
    om = np.array([u.mean(axis=1) for u in reps]); ov = np.array([u.var(axis=1) for u in reps])
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=300, nsamp_per_chain=60, n_chains=250, seed=77, want_u=True, **hm)
    mdl.close()
    U = out["u"].reshape(Q, 61, 250, order="F")[:, 1:, :]                  # [q, draw, chain]
    groups = U.reshape(Q, 60, 10, 25)                                      # 10 groups of 25 chains
    gm = groups.mean(axis=(1, 3)).T; gv = groups.var(axis=(1, 3)).T        # (10, Q)
    for name, a, b in (("mean", gm, om), ("variance", gv, ov)):
        se = np.sqrt(a.var(axis=0, ddof=1) / 10 + b.var(axis=0, ddof=1) / 10)
        z = np.abs(a.mean(axis=0) - b.mean(axis=0)) / se
        assert np.max(z) < 4.5, (family, name, float(np.max(z)))
    assert abs(out["stats"]["accept_rate"] - 0.9) < 0.08


def test_simlik_importance_forms_agree_while_exp_is_in_range(gctx):
    """likelihood.h:101-105 writes the importance-weighted objective as -log(exp(ll + logl) / exp(denomD)); the library evaluates it in log
    space by default and as written behind gmb_mcml_set_importance_form(1).  On a small model (|ll + logl| < 745) both give the same fit."""
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = synth.config2(m=300, seed=11, ncl=8, nt=4, nind=6)
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    a = (cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["U"], cfg["family"], cfg["link"], start)
    f0 = g.mcml_simlik(*a)
    try:
        g.mcml_set_importance_form(True)
        f1 = g.mcml_simlik(*a)
    finally:
        g.mcml_set_importance_form(False)
    assert np.max(np.abs(f0["beta"] - f1["beta"])) <= 1e-5 and np.max(np.abs(f0["theta"] - f1["theta"])) <= 1e-5
    # a model whose log-likelihood is below -745
    big = synth.config2(m=64, seed=12, ncl=40, nt=5, nind=10)              # n = 2000
    sb = np.concatenate([big["beta"] * 0.9, big["theta"], [1.0]])
    ab = (big["cov"], big["data"], big["eff_range"], big["Z"], big["X"], big["y"], big["U"], big["family"], big["link"], sb)
    ok = g.mcml_simlik(*ab)
    assert np.max(np.abs(ok["beta"] - sb[:big["P"]])) > 1e-3               # log-space form: the optimiser moves
    try:
        g.mcml_set_importance_form(True)
        with pytest.raises(g.GmbError) as e:                               # reference form: exp(ll + logl) underflows, the objective is not finite anywhere
            g.mcml_simlik(*ab)
        assert "not finite" in str(e.value)
    finally:
        g.mcml_set_importance_form(False)
