"""Family/link codes beyond the north-star's three: poisson/identity (2), binomial/log (4), binomial/identity (5),
binomial/probit (6), gaussian/log (8) — log-likelihood (moremaths.h:41-82), MCNR sums (mcmloptim.h:198-236 with the reconstructed
dhdmu of App. C.3), log_prob / log_grad (mcmlmodel.h:138-279) and the chain against the oracle, which restates the reference formulas
as written (including the double logarithm of gaussian/log and the sign of the binomial/log gradient)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {2: ("poisson", "identity"), 4: ("binomial", "log"), 5: ("binomial", "identity"), 6: ("binomial", "probit"), 8: ("gaussian", "log")}


def make(fl, n=230, P=3, Q=24, m=300, seed=5):
    rng = np.random.default_rng(seed + fl)
    X = np.column_stack([np.ones(n), 0.1 * rng.standard_normal((n, P - 1))])
    g = rng.integers(0, Q, n)
    Z = np.zeros((n, Q)); Z[np.arange(n), g] = 1.0
    if fl % 2 == 0 and fl != 6:
        Z = Z + 0.05 * rng.standard_normal((n, Q))              # a dense Z L for the even cases: the n x Q contraction path
    sd = {2: 0.15, 4: 0.1, 5: 0.03, 6: 0.5, 8: 0.1}[fl]
    L = np.linalg.cholesky(sd ** 2 * (0.6 * np.eye(Q) + 0.4 * np.ones((Q, Q)) / Q))
    beta = np.array([{2: 4.0, 4: -1.2, 5: 0.5, 6: 0.2, 8: 0.8}[fl], 0.3, -0.2])[:P]
    U = np.asfortranarray(L @ rng.standard_normal((Q, m)))
    eta = X @ beta + Z @ (L @ rng.standard_normal(Q))
    if fl == 2:
        y = rng.poisson(eta).astype(float)
    elif fl in (4, 5, 6):
        y = (rng.random(n) < 0.5).astype(float)
    else:
        y = np.exp(np.exp(0.3 * rng.standard_normal(n) + 0.5) + 1.0)     # log(log y) is finite
    return dict(X=np.asfortranarray(X), Z=np.asfortranarray(Z), y=y, L=np.asfortranarray(L), beta=beta, U=U, n=n, P=P, Q=Q, m=m)


@pytest.mark.parametrize("fl", sorted(CASES))
def test_family_link_codes_against_the_oracle(gctx, oracle, fl):
    import glmmrmcml_b200 as g
    fam, link = CASES[fl]
    assert oracle.flink(fam, link) == fl
    c = make(fl)
    sig = 0.7 if fl == 8 else 1.0
    y_model = np.log(c["y"]) if fl == 8 else c["y"]              # what mcmlModel holds after its constructor (mcmlmodel.h:90-92)
    mdl = g.Model(gctx, c["X"], c["Z"], c["y"], fam, link)
    assert mdl.flink == fl
    mdl.set_u(c["U"])
    zd = oracle.gemm(c["Z"], c["U"]); xb = c["X"] @ c["beta"]
    ref = oracle.loglik_zd(zd, xb, y_model, sig, fl)
    assert np.isfinite(ref)
    got = mdl.log_likelihood(c["beta"], sig)
    assert abs(got - ref) <= 1e-10 * abs(ref), (got, ref)
    B = np.asfortranarray(c["beta"][:, None] * (1 + 1e-3 * np.arange(9))[None, :])
    llb = mdl.log_likelihood_batch(B, np.full(9, sig))
    for k in (0, 4, 8):
        r = oracle.loglik_zd(zd, c["X"] @ B[:, k], y_model, sig, fl)
        assert abs(llb[k] - r) <= 1e-10 * abs(r)
    if fl != 6:
        nr = mdl.mcnr(c["beta"], sig)
        rr = oracle.mcnr(c["X"], c["Z"], c["U"], y_model, c["beta"], sig, fl)
        assert rr["rc"] == 0
        sc = np.max(np.abs(rr["xtwx"]))
        assert np.max(np.abs(nr["xtwx"] - rr["xtwx"])) <= 1e-10 * sc
        assert np.max(np.abs(nr["score"] - rr["score"])) <= 1e-10 * max(sc, np.max(np.abs(rr["score"])))
        assert abs(nr["sigma"] - rr["sigma"]) <= 1e-10 * rr["sigma"]
        assert np.max(np.abs(nr["beta_incr"] - rr["beta_incr"])) <= 1e-8 * max(1.0, np.max(np.abs(rr["beta_incr"])))
    else:
        with pytest.raises(g.GmbError) as e:
            mdl.mcnr(c["beta"], sig)
        assert e.value.code == 2                                  # GMB_EFAMILY: glmmrBase's probit dhdmu is not in the reference tree
    # target density, gradient and a short chain (two-contraction sampler; the structure-aware and on-chip kernels serve codes 1, 3, 7)
    ZL = c["Z"] @ c["L"]
    rng = np.random.default_rng(1)
    V = np.asfortranarray(0.2 * rng.standard_normal((c["Q"], 4)))
    lp, gr = mdl.log_prob_grad(c["L"], c["beta"], sig, V)
    for k in range(4):
        lr = oracle.log_prob(ZL, xb, y_model, sig, fl, V[:, k]); gg = oracle.log_grad(ZL, xb, y_model, sig, fl, V[:, k])
        assert abs(lp[k] - lr) <= 1e-10 * abs(lr)
        assert np.max(np.abs(gr[:, k] - gg)) <= 1e-10 * np.max(np.abs(gg))
    out = mdl.hmc_sample(c["L"], c["beta"], sig, warmup=6, nsamp_per_chain=3, lam=0.02, max_steps=8, target_accept=0.9, n_chains=5, seed=9, want_v=True)
    assert out["stats"]["kernel_variant"] == 1
    for ch in (0, 4):
        rc = oracle.hmc_chain(ZL, c["L"], xb, y_model, sig, fl, 6, 3, 0.02, 8, 0.9, 9, chain=ch)
        assert np.max(np.abs(out["v"][:, ch * 4:(ch + 1) * 4] - rc["v"])) <= 1e-7
    mdl.close()


def test_unsupported_codes_fail_loudly(gctx):
    import glmmrmcml_b200 as g
    c = make(2)
    for fam, link in (("Gamma", "log"), ("beta", "logit"), ("binomial", "cloglog")):
        with pytest.raises(g.GmbError) as e:
            g.Model(gctx, c["X"], c["Z"], c["y"], fam, link)
        assert e.value.code == 2
