"""BASELINE.json's configurations at their FULL stated size.

C2 (n = 500, Q = 50, m = 10^4): the oracle checks every E-step quantity directly (it needs milliseconds at this size), plus
size-independent properties — additivity over column splits, determinism, agreement of the sampler's kernel variants, Monte-Carlo
agreement across seeds.  C3 (n = Q = 10^4), C4 (m = 10^5) and C5 (n = 5 10^4, Q = 5 10^3, m = 10^4): tools/bench_configs.run_config,
the block bench.py emits, whose `parity` object compares the GPU with the oracle (LAPACK for the one dense 10^4 block) on the same inputs."""
import numpy as np
import pytest

from glmmrmcml_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(gctx):
    import glmmrmcml_b200 as g
    cfg = synth.config2(m=10_000)
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    cv = g.Covariance(gctx, cfg["cov"], cfg["data"], cfg["eff_range"])
    yield cfg, mdl, cv
    mdl.close(); cv.close()


def test_estep_is_additive_over_column_splits(c2):
    cfg, mdl, cv = c2
    U = cfg["U"]; m = U.shape[1]; cut = 3777
    beta = cfg["beta"] * 1.05
    vals = {}
    for name, cols in (("all", slice(0, m)), ("a", slice(0, cut)), ("b", slice(cut, m))):
        Us = np.asfortranarray(U[:, cols])
        mdl.set_u(Us)
        nr = mdl.mcnr(beta, 1.0)
        vals[name] = (mdl.log_likelihood(beta, 1.0), nr["xtwx"], nr["score"], nr["sigma"], cv.loglik(cfg["theta"], Us), Us.shape[1])
    la, lb, lall = vals["a"], vals["b"], vals["all"]
    wa, wb = la[5] / m, lb[5] / m
    assert abs(wa * la[0] + wb * lb[0] - lall[0]) <= 1e-12 * abs(lall[0])
    assert np.max(np.abs(wa * la[1] + wb * lb[1] - lall[1])) <= 1e-12 * np.max(np.abs(lall[1]))
    assert np.max(np.abs(wa * la[2] + wb * lb[2] - lall[2])) <= 1e-11 * max(1.0, np.max(np.abs(lall[2])))
    assert abs(wa * la[3] + wb * lb[3] - lall[3]) <= 1e-12 * lall[3]
    assert abs(wa * la[4] + wb * lb[4] - lall[4]) <= 1e-12 * abs(lall[4])


def test_estep_stream_larger_than_l2_is_additive_and_deterministic(gctx):
    """1 GB of zd (the roofline probe's size): two halves add up to the whole, repeated evaluations are bitwise equal."""
    import glmmrmcml_b200 as g
    cfg = synth.config2(m=64)
    rng = np.random.default_rng(7)
    m = 250_000
    U = np.asfortranarray(cfg["L"] @ rng.standard_normal((cfg["Q"], m)))
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    mdl.set_u(U)
    full = mdl.log_likelihood(cfg["beta"], 1.0)
    assert mdl.log_likelihood(cfg["beta"], 1.0) == full
    mdl.set_u(np.asfortranarray(U[:, :100_000])); a = mdl.log_likelihood(cfg["beta"], 1.0)
    mdl.set_u(np.asfortranarray(U[:, 100_000:])); b = mdl.log_likelihood(cfg["beta"], 1.0)
    assert abs(0.4 * a + 0.6 * b - full) <= 1e-12 * abs(full)
    mdl.close()


def test_sampler_cluster_variants_agree_at_bench_size(c2):
    """250 chains, the bench's proposal settings (shortened): cluster sizes 1, 2 and 4 follow the same chains (summation order of the
    gradient differs, so states agree to rounding amplified by the leapfrog map, and the accept decisions are identical)."""
    import glmmrmcml_b200 as g
    cfg, mdl, cv = c2
    outs = {}
    try:
        g.hmc_set_variant(2)                       # the dense on-chip kernel (the dispatcher alone picks the structure-aware one for C2)
        for cs in (1, 2, 4):
            g.hmc_set_cluster_size(cs)
            outs[cs] = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=30, nsamp_per_chain=3, lam=5.0, max_steps=100, target_accept=0.95,
                                      n_chains=250, seed=20221208, want_u=False, want_v=True)
            assert outs[cs]["stats"]["kernel_variant"] == 2
    finally:
        g.hmc_set_cluster_size(0)
        g.hmc_set_variant(0)
    for cs in (2, 4):
        assert np.max(np.abs(outs[cs]["v"] - outs[1]["v"])) <= 1e-8
        assert outs[cs]["stats"]["accept_rate"] == outs[1]["stats"]["accept_rate"]
    # the structure-aware kernel (sparse Z L) follows the same chains
    sp = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=30, nsamp_per_chain=3, lam=5.0, max_steps=100, target_accept=0.95,
                        n_chains=250, seed=20221208, want_u=False, want_v=True)
    assert sp["stats"]["kernel_variant"] == 3
    assert np.max(np.abs(sp["v"] - outs[1]["v"])) <= 1e-8
    assert sp["stats"]["accept_rate"] == outs[1]["stats"]["accept_rate"]
    outs[3] = sp
    # same seed, same variant: bitwise reproducible
    again = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=30, nsamp_per_chain=3, lam=5.0, max_steps=100, target_accept=0.95,
                           n_chains=250, seed=20221208, want_u=False, want_v=True)
    assert np.array_equal(again["v"], outs[3]["v"])


def test_sampler_posterior_means_agree_across_seeds_at_bench_size(c2):
    """Full bench draw (250 chains x 40 columns after a 500-iteration warm-up): the posterior mean of every random effect from two
    seeds agrees within 6 Monte-Carlo standard errors (chains are independent, so the error comes from per-chain means)."""
    cfg, mdl, cv = c2
    means, ses = [], []
    for seed in (5, 6):
        out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=500, nsamp_per_chain=39, lam=5.0, max_steps=100, target_accept=0.95,
                             n_chains=250, seed=seed, want_u=True)
        assert 0.9 < out["stats"]["accept_rate"] <= 1.0
        Uc = out["u"].reshape(cfg["Q"], 40, 250, order="F").transpose(0, 2, 1)[:, :, 1:]    # columns are chain-major: [q, chain, draw]
        cm = Uc.mean(axis=2)
        means.append(cm.mean(axis=1)); ses.append(cm.std(axis=1, ddof=1) / np.sqrt(250))
    z = np.abs(means[0] - means[1]) / np.sqrt(ses[0] ** 2 + ses[1] ** 2)
    assert np.max(z) < 6.0, np.max(z)


def test_c2_full_size_against_the_oracle(c2, oracle):
    """m = 10^4: log-likelihood (single, batched), MCNR sums and mvn_ll against the oracle at 1e-10."""
    cfg, mdl, cv = c2
    fl = oracle.flink(cfg["family"], cfg["link"])
    mdl.set_u(cfg["U"])
    zd = oracle.gemm(cfg["Z"], cfg["U"])
    rng = np.random.default_rng(2)
    B = np.asfortranarray(cfg["beta"][:, None] + 1e-2 * rng.standard_normal((cfg["P"], 16)))
    llb = mdl.log_likelihood_batch(B, np.ones(16))
    for k in range(16):
        ref = oracle.loglik_zd(zd, cfg["X"] @ B[:, k], cfg["y"], 1.0, fl)
        assert abs(llb[k] - ref) <= 1e-10 * abs(ref)
        if k < 3:
            assert abs(mdl.log_likelihood(B[:, k], 1.0) - ref) <= 1e-10 * abs(ref)
    nr = mdl.mcnr(cfg["beta"], 1.0)
    ref = oracle.mcnr(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"], 1.0, fl)
    assert np.max(np.abs(nr["xtwx"] - ref["xtwx"])) <= 1e-10 * np.max(np.abs(ref["xtwx"]))
    assert np.max(np.abs(nr["score"] - ref["score"])) <= 1e-10 * np.max(np.abs(ref["xtwx"]))
    assert abs(nr["sigma"] - ref["sigma"]) <= 1e-10 * ref["sigma"]
    for th in (cfg["theta"], cfg["theta"] * np.array([1.3, 0.9])):
        r = oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], th, cfg["U"])
        assert abs(cv.loglik_model(th, mdl) - r) <= 1e-10 * abs(r)
        assert abs(cv.loglik(th, cfg["U"]) - r) <= 1e-10 * abs(r)


@pytest.mark.parametrize("name", ["C3", "C4", "C5"])
def test_large_configs_full_size_parity(gctx, oracle, name):
    """C3 at n = Q = 10^4 (Cholesky + solves of one 10^4 block, factored sampler), C4 at m = 10^5 (8 GB of samples, component sampler),
    C5 at n = 5 10^4, Q = 5 10^3, m = 10^4 — every parity field of the bench block within its tolerance."""
    import glmmrmcml_b200 as g
    from tools import bench_configs as bc
    out = bc.run_config(name, bc.Env(g, gctx), size="full", oracle=oracle)
    p = out["parity"]
    assert p["loglik"]["rel_err"] <= 1e-10 and p["loglik"]["rel_err_default_path"] <= 1e-10, p
    assert max(p["mcnr"]["xtwx_rel_err"], p["mcnr"]["score_rel_err"], p["mcnr"]["sigma_rel_err"]) <= 1e-10, p
    assert p["mvn_ll"]["rel_err"] <= 1e-10, p
    assert p["loglik_on_sampled_u"]["rel_err"] <= 1e-10, p
    assert p["log_prob"]["rel_err"] <= 1e-10 and p["log_grad"]["rel_err"] <= 1e-10, p
    assert p["chain"]["max_abs_err"] <= 1e-7 and p["chain"]["same_kernel_as_timed_run"], p
    if "chol" in p:
        assert p["chol"]["ok"], p
    assert p["ok"]
