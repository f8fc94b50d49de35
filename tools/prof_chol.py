"""Profiling driver: one large-block factorisation (+ one mvn_ll) at n = NLOC, for `ncu --metrics gpu__time_duration.sum`."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(1)
xy = rng.random((nloc, 2))
cov = np.array([[0, nloc, 13, 2, 0]], dtype=np.int32); data = np.concatenate([xy[:, 0], xy[:, 1]])
ctx = g.Context(0)
cv = g.Covariance(ctx, cov, data, np.zeros(1))
th = np.array([0.25, 0.1])
cv.logdet(th)
ctx.sync(); ctx.timer_start(); cv.logdet(th * 1.001); print("factor ms", ctx.timer_stop())
if m:
    U = np.asfortranarray(rng.standard_normal((nloc, m)))
    Z = np.zeros((8, nloc), order="F"); Z[np.arange(8), np.arange(8)] = 1.0
    mdl = g.Model(ctx, np.ones((8, 1), order="F"), Z, np.zeros(8), "gaussian", "identity")
    mdl.set_u(U)                                     # samples resident on the device
    cv.loglik_model(th * 1.002, mdl)
    ctx.sync(); ctx.timer_start(); v = cv.loglik_model(th * 1.003, mdl); t = ctx.timer_stop()
    print("mvn_ll (factor + solve of %d device-resident columns) ms" % m, t)
    ctx.timer_start(); v = cv.loglik_model(th * 1.003, mdl); t2 = ctx.timer_stop()
    print("second evaluation at the same theta (factor cached; default path: Gram factor when m >= 2 n) ms", t2)
    g.cov_set_gram(False)
    cv.loglik_model(th * 1.004, mdl)
    ctx.sync(); ctx.timer_start(); cv.loglik_model(th * 1.004, mdl); t3 = ctx.timer_stop()
    g.cov_set_gram(True)
    print("streaming forward substitution of the %d columns (factor cached) ms" % m, t3, " -> %.1f TFLOP/s" % (nloc * float(nloc) * m / (t3 * 1e-3) / 1e12))
