"""GPU parity of the covariance path (K4 D(theta) build + Cholesky, K5 mvn_ll) against the CPU oracle."""
import numpy as np
import pytest

from glmmrmcml_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def mixed_blocks(rng):
    """blocks of many sizes and kernels in one specification: gr 1x1, gr*ar1 5x5, fexp 12x12 / 40x40 / 150x150, sqexp 7x7"""
    cov, data = [], []
    b = 0
    for _ in range(5):
        cov.append([b, 1, 1, 1, 0]); data += [float(b + 1)]; b += 1
    for c in range(4):
        cov.append([b, 5, 1, 1, 0]); cov.append([b, 5, 3, 1, 1]); data += [c + 1.0] * 5 + [1.0, 2.0, 3.0, 4.0, 5.0]; b += 1
    for nb in (12, 40, 150):
        xy = rng.random((nb, 2))
        cov.append([b, nb, 13, 2, 2]); data += list(xy[:, 0]) + list(xy[:, 1]); b += 1
    x = rng.random(7)
    cov.append([b, 7, 4, 1, 4]); data += list(x); b += 1
    theta = np.array([0.3, 0.7, 0.25, 0.15, 0.5, 0.4])
    return np.array(cov, dtype=np.int32), np.array(data), theta


CASES = {
    "C1": lambda: synth.config1(m=250),
    "C2": lambda: synth.config2(m=1000),
    "C3": lambda: synth.config3(nloc=250, m=250),
    "C4": lambda: synth.config4(ncl=100, nt=10, k=1, m=300),
    "C5": lambda: synth.config5(nloc=700, nobs=2, m=200),
    # one dense block spanning several outer blocks (256) and trailing-update block columns (512) of the two-level Cholesky
    "C3-1100": lambda: synth.config3(nloc=1100, m=130),
}


@pytest.mark.parametrize("name", list(CASES))
def test_chol_and_mvn_ll(name, gctx, oracle):
    import glmmrmcml_b200 as g
    cfg = CASES[name]()
    cv = g.Covariance(gctx, cfg["cov"], cfg["data"], cfg["eff_range"])
    assert (cv.B, cv.Q, cv.R) == oracle.cov_dims(cfg["cov"], cfg["data"])
    for scale in (1.0, 1.3):
        theta = cfg["theta"] * np.array([scale, 1.0 / scale if name in ("C3", "C5", "C3-1100") else min(0.95, scale * cfg["theta"][1]) / cfg["theta"][1]])[: cfg["theta"].size]
        Lw = oracle.genD(cfg["cov"], cfg["data"], cfg["eff_range"], theta, chol=True)
        Lg = cv.genD(theta, chol=True)
        assert np.max(np.abs(Lg - Lw)) <= 1e-11 * np.max(np.abs(Lw)) * max(1.0, np.linalg.cond(Lw))
        Dw = oracle.genD(cfg["cov"], cfg["data"], cfg["eff_range"], theta, chol=False)
        Dg = cv.genD(theta, chol=False)
        assert np.max(np.abs(Dg - Dw)) <= 1e-14 * np.max(np.abs(Dw))
        want = oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], theta, cfg["U"])
        got = cv.loglik(theta, cfg["U"])
        kappa = np.linalg.cond(Lw) ** 2
        assert abs(got - want) <= max(RTOL, 1e-15 * kappa) * abs(want), (got, want, kappa)
        assert abs(cv.logdet(theta) - oracle.logdet(cfg["cov"], cfg["data"], cfg["eff_range"], theta)) <= 1e-10 * max(1.0, abs(want))
    cv.close()


def test_mixed_blocks(gctx, oracle):
    import glmmrmcml_b200 as g
    rng = np.random.default_rng(5)
    cov, data, theta = mixed_blocks(rng)
    eff = np.zeros(cov.shape[0])
    cv = g.Covariance(gctx, cov, data, eff)
    L = oracle.genD(cov, data, eff, theta, chol=True)
    U = np.asfortranarray(L @ rng.standard_normal((L.shape[0], 333)))
    assert np.max(np.abs(cv.genD(theta) - L)) <= 1e-10 * np.max(np.abs(L))
    want = oracle.mvn_loglik(cov, data, eff, theta, U)
    got = cv.loglik(theta, U)
    assert abs(got - want) <= 1e-9 * abs(want), (got, want)
    # the stateless reference-named entry point gives the same number
    assert g.mvn_ll(cov, data, eff, theta, U) == got
    cv.close()


def test_vector_u_and_single_column(gctx, oracle):
    """mvn_ll accepts a vector u (man/mvn_ll.Rd:19 'Matrix (or vector)')."""
    import glmmrmcml_b200 as g
    cfg = synth.config2(m=1)
    want = oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], cfg["U"])
    got = g.mvn_ll(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], cfg["U"][:, 0])
    assert abs(got - want) <= RTOL * abs(want)


def test_not_positive_definite(gctx):
    import glmmrmcml_b200 as g
    cfg = synth.config2(m=4)
    cv = g.Covariance(gctx, cfg["cov"], cfg["data"], cfg["eff_range"])
    with pytest.raises(g.GmbError) as e:
        cv.loglik(np.array([0.25, 1.5]), cfg["U"])        # ar1 with rho > 1 is not positive definite
    assert e.value.code == 4
    with pytest.raises(g.GmbError) as e:
        g.Covariance(gctx, np.array([[0, 3, 5, 1, 0]], dtype=np.int32), np.zeros(3), None)   # matern: not supported
    assert e.value.code == 7
    cv.close()


@pytest.mark.parametrize("name", ["C1", "C2", "C4"])
def test_gram_path_equals_stream_and_oracle(name, gctx, oracle):
    """mvn_ll on a model's device-resident samples: the Gram-matrix evaluation (default for blocks <= 16) and the streaming evaluation
    agree with the oracle at several theta on the same samples; the Gram matrices are rebuilt when the samples change."""
    import glmmrmcml_b200 as g
    cfg = CASES[name]()
    cv = g.Covariance(gctx, cfg["cov"], cfg["data"], cfg["eff_range"])
    mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    rng = np.random.default_rng(4)
    for rep in range(2):
        U = np.asfortranarray(cfg["U"] * (1.0 + 0.3 * rep) + 0.01 * rng.standard_normal(cfg["U"].shape))
        mdl.set_u(U)
        for scale in (1.0, 0.8, 1.15):
            theta = cfg["theta"] * np.array([scale, min(0.95, scale * cfg["theta"][1]) / cfg["theta"][1]])[: cfg["theta"].size]
            want = oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], theta, U)
            got = cv.loglik_model(theta, mdl)
            try:
                g.cov_set_gram(False); stream = cv.loglik_model(theta, mdl)
            finally:
                g.cov_set_gram(True)
            assert abs(stream - want) <= 1e-10 * abs(want)
            assert abs(got - want) <= 1e-10 * abs(want), (got, stream, want)
            assert abs(cv.logdet(theta) - oracle.logdet(cfg["cov"], cfg["data"], cfg["eff_range"], theta)) <= 1e-10 * max(1.0, abs(want))
    # batched evaluation (one launch for all points): bitwise the single evaluations; -inf where D(theta) is not positive definite
    R = cfg["theta"].size
    pts = np.asfortranarray(np.stack([cfg["theta"] * s for s in (1.0, 0.7, 0.9, 1.1, 1.0)], axis=1))
    if name != "C1":
        pts[:, 2] = [0.3, 1.2]                                   # ar1 parameter > 1
    single = []
    for k in range(pts.shape[1]):
        try:
            single.append(cv.loglik_model(pts[:, k], mdl))
        except g.GmbError:
            single.append(-np.inf)
    batch = cv.loglik_model_batch(pts, mdl)
    assert np.array_equal(batch, np.array(single))
    try:
        g.cov_set_gram(False)
        assert np.allclose(cv.loglik_model_batch(pts, mdl), batch, rtol=1e-10, atol=0)     # falls back to single streaming evaluations
    finally:
        g.cov_set_gram(True)
    if name != "C1":
        assert batch[2] == -np.inf
        with pytest.raises(g.GmbError):
            cv.loglik_model(np.array([0.3, 1.2]), mdl)          # ar1 parameter > 1: not positive definite
    mdl.close(); cv.close()


@pytest.mark.parametrize("fid", [7, 8, 9])
def test_wendland_compact_support_blocks(gctx, oracle, fid):
    """Covariance function ids 7-9 (wend0 / wend1 / wend2, compact support through eff_range; SURVEY App. C.2 reconstruction, N4): D(theta), its
    Cholesky factor and mvn_ll against the oracle, small (warp) and large (blocked) blocks."""
    import glmmrmcml_b200 as g
    rng = np.random.default_rng(40 + fid)
    for nloc, m in ((12, 64), (150, 40)):
        xy = rng.random((nloc, 2))
        cov = np.array([[0, nloc, fid, 2, 0]], dtype=np.int32)
        data = np.concatenate([xy[:, 0], xy[:, 1]])
        eff = np.array([0.45])
        theta = np.array([0.8, 3.0 + fid])                      # smoothness large enough for positive definiteness in 2-D
        cv = g.Covariance(gctx, cov, data, eff)
        D = cv.genD(theta, chol=False); L = cv.genD(theta, chol=True)
        Do = oracle.genD(cov, data, eff, theta, chol=False); Lo = oracle.genD(cov, data, eff, theta, chol=True)
        assert np.count_nonzero(Do == 0.0) > 0                  # compact support: exact zeros beyond eff_range
        assert np.max(np.abs(D - Do)) <= 1e-14 and np.max(np.abs(L - Lo)) <= 1e-11
        U = np.asfortranarray(Lo @ rng.standard_normal((nloc, m)))
        r = oracle.mvn_loglik(cov, data, eff, theta, U)
        assert abs(cv.loglik(theta, U) - r) <= 1e-10 * abs(r)
        cv.close()
    with pytest.raises(g.GmbError):                             # compact support needs its range
        g.Covariance(gctx, cov, data, np.array([0.0]))


def test_identical_blocks_are_factorised_once(gctx, oracle):
    """gr(cl)*ar1(t) repeats ONE block per cluster (mcmldmatrix.h:26-36 loops over all of them).  The Gram path finds the classes
    (here: 60 identical 10 x 10 blocks, 7 identical 4 x 4 blocks with another parameter set, 3 singletons), factorises one block per class on the
    summed Gram matrix and agrees with the per-block evaluation and with the oracle."""
    import glmmrmcml_b200 as g
    rng = np.random.default_rng(11)
    cov, data, b = [], [], 0
    for c in range(60):                                   # gr(cl) * ar1(t), label differs per cluster
        cov += [[b, 10, 1, 1, 0], [b, 10, 3, 1, 1]]; data += [c + 1.0] * 10 + list(np.arange(1.0, 11.0)); b += 1
    for c in range(7):                                    # gr(cl2) * fexp0(x) on a shared design of 4 points
        cov += [[b, 4, 1, 1, 2], [b, 4, 2, 1, 3]]; data += [100.0 + c] * 4 + [0.0, 0.5, 1.25, 2.0]; b += 1
    for c in range(3):                                    # singletons: different point sets
        x = rng.random(6)
        cov += [[b, 6, 2, 1, 4]]; data += list(x); b += 1
    cov = np.array(cov, dtype=np.int32); data = np.array(data); eff = np.zeros(len(cov))
    theta = np.array([0.4, 0.6, 0.5, 0.8, 0.7])
    cv = g.Covariance(gctx, cov, data, eff)
    assert cv.B == 70 and cv.block_classes == 5
    Q = cv.Q
    n = 40
    Z = np.zeros((n, Q), order="F"); Z[np.arange(n), rng.integers(0, Q, n)] = 1.0
    mdl = g.Model(gctx, np.ones((n, 1), order="F"), Z, rng.poisson(2.0, n).astype(float), "poisson", "log")
    L = cv.genD(theta, chol=True)
    U = np.asfortranarray(L @ rng.standard_normal((Q, 777)))
    mdl.set_u(U)
    pts = np.asfortranarray(np.stack([theta * s for s in (1.0, 0.8, 1.1)], axis=1))
    for k in range(3):
        want = oracle.mvn_loglik(cov, data, eff, pts[:, k], U)
        got = cv.loglik_model(pts[:, k], mdl)
        try:
            g.cov_set_block_classes(False); per_block = cv.loglik_model(pts[:, k], mdl)
        finally:
            g.cov_set_block_classes(True)
        assert abs(got - want) <= 1e-10 * abs(want) and abs(per_block - want) <= 1e-10 * abs(want), (got, per_block, want)
        assert abs(got - per_block) <= 1e-13 * abs(want)
        assert abs(cv.logdet(pts[:, k]) - oracle.logdet(cov, data, eff, pts[:, k])) <= 1e-10 * abs(want)    # per-block factor cache still right
    batch = cv.loglik_model_batch(pts, mdl)
    assert np.array_equal(batch, np.array([cv.loglik_model(pts[:, k], mdl) for k in range(3)]))
    bad = theta.copy(); bad[1] = 1.3                       # ar1 parameter > 1: the class's block is not positive definite
    with pytest.raises(g.GmbError):
        cv.loglik_model(bad, mdl)
    # a second covariance object whose blocks all differ keeps the per-block path
    cfg = synth.config3(nloc=12, m=10)
    cv2 = g.Covariance(gctx, np.array([[0, 5, 13, 1, 0], [1, 5, 13, 1, 0]], dtype=np.int32), np.concatenate([rng.random(5), rng.random(5)]), np.zeros(2))
    assert cv2.block_classes == 2
    cv2.close(); mdl.close(); cv.close()


def test_large_blocks_through_the_cholesky_factor_of_their_gram_matrix(gctx, oracle):
    """mvn_ll on a model's samples, blocks beyond the warp-sized paths: with m >= 2 n_b samples on this rank the library keeps
    C_b = chol(U_b U_b^T) per sample matrix and evaluates sum_j ||L_b^-1 u_bj||^2 = ||L_b^-1 C_b||_F^2 (n_b^3 / 3 flop per theta instead of
    n_b^2 m).  Against the streaming evaluation and the oracle at several theta; the fall-backs (too few samples, singular Gram matrix)."""
    import glmmrmcml_b200 as g
    rng = np.random.default_rng(17)
    for nloc, m in ((150, 400), (700, 1500), (1100, 2300)):
        cfg = synth.config3(nloc=nloc, m=m)
        cv = g.Covariance(gctx, cfg["cov"], cfg["data"], cfg["eff_range"])
        mdl = g.Model(gctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
        for rep in range(2):                                   # the factor of the Gram matrix follows the samples
            U = np.asfortranarray(cfg["U"] * (1.0 + 0.2 * rep) + 0.01 * rng.standard_normal(cfg["U"].shape))
            mdl.set_u(U)
            for scale in (1.0, 0.85, 1.1):
                theta = cfg["theta"] * scale
                want = oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], theta, U)
                got = cv.loglik_model(theta, mdl)
                try:
                    g.cov_set_gram(False); stream = cv.loglik_model(theta, mdl)
                finally:
                    g.cov_set_gram(True)
                assert abs(stream - want) <= 1e-10 * abs(want), (nloc, stream, want)
                assert abs(got - want) <= 1e-10 * abs(want), (nloc, got, stream, want)
        # fewer than 2 n samples: the samples are streamed (same value as with the Gram path off, bit for bit)
        mdl.set_u(np.asfortranarray(U[:, : nloc + 7]))
        a = cv.loglik_model(cfg["theta"], mdl)
        try:
            g.cov_set_gram(False); b = cv.loglik_model(cfg["theta"], mdl)
        finally:
            g.cov_set_gram(True)
        assert a == b
        # a singular Gram matrix (2 n + 10 columns spanning 40 dimensions): detected, streamed
        Us = np.asfortranarray(U[:, :40] @ rng.standard_normal((40, 2 * nloc + 10)))
        mdl.set_u(Us)
        want = oracle.mvn_loglik(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], Us)
        assert abs(cv.loglik_model(cfg["theta"], mdl) - want) <= 1e-10 * abs(want)
        mdl.close(); cv.close()


def test_gram_factor_path_with_several_large_blocks(gctx, oracle):
    """two large blocks (Gram-factor path each), one medium and a few small ones in one covariance; and an odd block offset (unaligned rows of U:
    that block streams its samples)."""
    import glmmrmcml_b200 as g
    rng = np.random.default_rng(23)
    for sizes in ((100, 130, 40, 6), (7, 101, 90)):            # second layout: the 101-block starts at row 7 (odd)
        cov, data, b = [], [], 0
        for nb in sizes:
            xy = rng.random((nb, 2))
            cov.append([b, nb, 13, 2, 0]); data += list(xy[:, 0]) + list(xy[:, 1]); b += 1
        cov = np.array(cov, dtype=np.int32); data = np.array(data); eff = np.zeros(len(cov))
        theta = np.array([0.3, 0.2])
        cv = g.Covariance(gctx, cov, data, eff)
        Q, n = cv.Q, 30
        Z = np.zeros((n, Q), order="F"); Z[np.arange(n), rng.integers(0, Q, n)] = 1.0
        mdl = g.Model(gctx, np.ones((n, 1), order="F"), Z, rng.standard_normal(n), "gaussian", "identity")
        U = np.asfortranarray(cv.genD(theta, chol=True) @ rng.standard_normal((Q, 300)))
        mdl.set_u(U)
        for scale in (1.0, 1.2):
            want = oracle.mvn_loglik(cov, data, eff, theta * scale, U)
            got = cv.loglik_model(theta * scale, mdl)
            try:
                g.cov_set_gram(False); stream = cv.loglik_model(theta * scale, mdl)
            finally:
                g.cov_set_gram(True)
            assert abs(got - want) <= 1e-10 * abs(want) and abs(stream - want) <= 1e-10 * abs(want), (sizes, got, stream, want)
        mdl.close(); cv.close()
