"""fp32 mode (gmb_model_create_prec(..., 32); BASELINE.json north_star: "within 1e-10 relative in fp64 (1e-5 in fp32)"): the streamed E-step
matrices zd = Z u and F = exp(+-zd) are stored as float (half the HBM bytes per evaluation), all sums are accumulated in fp64.  Every E-step
quantity of the five configurations against the fp64 ORACLE at 1e-5 relative, the row-statistic and streaming paths, the dense and the
gathered zd build, and the fp64 mode of the same library as a second reference."""
import numpy as np
import pytest

from glmmrmcml_b200 import synth

pytestmark = pytest.mark.gpu
TOL32 = 1e-5

CASES = {
    "C1": lambda: synth.config1(m=250),                                   # Q = 60: dense zd contraction, narrowed
    "C2": lambda: synth.config2(m=2000),
    "C3": lambda: synth.config3(nloc=250, m=250),                         # gaussian: row statistics / stream; Z = I gathered
    "C4": lambda: synth.config4(ncl=100, nt=10, k=1, m=512),              # poisson, sparse Z gathered
    "C5": lambda: synth.config5(nloc=300, nobs=10, m=600),                # binomial, sparse Z gathered, factor matrix in float
    "ragged": lambda: synth.config4(ncl=37, nt=7, k=3, m=131),
}


@pytest.mark.parametrize("name", list(CASES) + ["C1-dense", "C2-dense"])
def test_fp32_mode_estep_against_the_fp64_oracle(gctx, oracle, name):
    import glmmrmcml_b200 as g
    dense = name.endswith("-dense")                # C1 / C2: E-step on every observation instead of on the 50 distinct rows
    name = name.replace("-dense", "")
    cfg = CASES[name]()
    fl = oracle.flink(cfg["family"], cfg["link"])
    m32 = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"], precision="fp32")
    m64 = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    g.estep_set_row_aggregation(not dense)
    try:
        m32.set_u(cfg["U"]); m64.set_u(cfg["U"])
    finally:
        g.estep_set_row_aggregation(True)
    assert m32.estep_rows() == (50 if (name in ("C1", "C2") and not dense) else cfg["n"])
    rng = np.random.default_rng(3)
    worst = 0.0
    for trial in range(3):
        beta = cfg["beta"] + 0.1 * trial * rng.standard_normal(cfg["P"])
        sigma = 1.0 + 0.3 * trial
        want = oracle.loglik_faithful(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], beta, sigma, fl)
        for stream_only in (False, True):                                  # default path (row statistics where they apply) and the streaming kernel
            g.estep_set_rowstats(not stream_only)
            try:
                got = m32.log_likelihood(beta, sigma)
            finally:
                g.estep_set_rowstats(True)
            assert abs(got - want) <= TOL32 * abs(want), (name, stream_only, got, want)
            worst = max(worst, abs(got - want) / abs(want))
        assert abs(m64.log_likelihood(beta, sigma) - want) <= 1e-10 * abs(want)
    B = np.asfortranarray(cfg["beta"][:, None] * (1 + 1e-3 * np.arange(12))[None, :])
    lb = m32.log_likelihood_batch(B, np.ones(12))
    lb64 = m64.log_likelihood_batch(B, np.ones(12))
    assert np.max(np.abs(lb - lb64) / np.abs(lb64)) <= TOL32
    nr = m32.mcnr(cfg["beta"], 1.0)
    ref = oracle.mcnr(cfg["X"], cfg["Z"], cfg["U"], cfg["y"], cfg["beta"], 1.0, fl)
    sc = np.max(np.abs(ref["xtwx"]))
    assert np.max(np.abs(nr["xtwx"] - ref["xtwx"])) <= TOL32 * sc
    assert np.max(np.abs(nr["score"] - ref["score"])) <= TOL32 * max(sc, np.max(np.abs(ref["score"])))
    assert abs(nr["sigma"] - ref["sigma"]) <= TOL32 * ref["sigma"]
    # the float rounding is really there (this is not the fp64 path under another name), and far inside the tolerance
    assert 1e-12 < worst < 1e-6, worst
    m32.close(); m64.close()


def test_fp32_mode_rejects_what_it_does_not_implement(gctx):
    import glmmrmcml_b200 as g
    cfg = synth.config2(m=8)
    with pytest.raises(g.GmbError) as e:
        g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "probit", precision="fp32")
    assert e.value.code == 2


@pytest.mark.parametrize("shape", [(700, 333, 500), (128, 64, 128), (130, 37, 259), (1500, 2049, 300)])
def test_tcgen05_3xtf32_contraction_of_a_dense_z(gctx, oracle, shape):
    """fp32 mode with a dense Z: zd = Z u runs on tcgen05.mma kind::tf32 with the 3xTF32 operand split, TMA staging and a TMEM accumulator
    (gemm_tf32.cu).  The log-likelihood (a sum over all n x m entries of zd) and the MCNR sums agree with the fp64 oracle at the fp32 mode's
    1e-5, with the same model run through the fp64 DMMA product narrowed to float, and plain TF32 would NOT (its 2^-11 operand rounding
    gives ~1e-4): the split is doing its job.  Shapes cover partial tiles in M, N and K."""
    import glmmrmcml_b200 as g
    n, Q, m = shape
    rng = np.random.default_rng(n + Q)
    X = np.asfortranarray(np.column_stack([np.ones(n), rng.standard_normal(n)]))
    Z = np.asfortranarray(rng.standard_normal((n, Q)) / np.sqrt(Q))
    beta = np.array([0.3, -0.2])
    U = np.asfortranarray(rng.standard_normal((Q, m)))
    y = (rng.random(n) < 0.5).astype(float)
    fl = oracle.flink("binomial", "logit")
    want = oracle.loglik_faithful(X, Z, U, y, beta, 1.0, fl)
    vals = {}
    for on in (True, False):
        g.estep_set_tf32(on)
        try:
            mdl = g.Model(gctx, X, Z, y, "binomial", "logit", precision="fp32")
            mdl.set_u(U)
            vals[on] = (mdl.log_likelihood(beta, 1.0), mdl.mcnr(beta, 1.0))
            mdl.close()
        finally:
            g.estep_set_tf32(True)
    for on in (True, False):
        assert abs(vals[on][0] - want) <= TOL32 * abs(want), (on, vals[on][0], want)
    ref = oracle.mcnr(X, Z, U, y, beta, 1.0, fl)
    assert np.max(np.abs(vals[True][1]["xtwx"] - ref["xtwx"])) <= TOL32 * np.max(np.abs(ref["xtwx"]))
    assert abs(vals[True][0] - vals[False][0]) <= 2e-6 * abs(want)
    assert vals[True][0] != vals[False][0]                              # two different contractions
