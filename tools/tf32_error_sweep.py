"""fp32 mode, dense Z: relative error of the log-likelihood (against the fp64 mode of the same library) as a function of the contraction length K = Q,
for the tcgen05 3xTF32 product and for the fp64 product narrowed to float."""
import sys, os, json, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
ctx = g.Context(0)
n, m = 1024, 512
rng = np.random.default_rng(1)
for Q in (256, 1024, 2048, 4096, 8192, 16384):
    X = np.asfortranarray(np.ones((n, 1))); Z = np.asfortranarray(rng.standard_normal((n, Q)) / np.sqrt(Q))
    U = np.asfortranarray(rng.standard_normal((Q, m))); y = (rng.random(n) < 0.5).astype(float); beta = np.array([0.1])
    m64 = g.Model(ctx, X, Z, y, "binomial", "logit"); m64.set_u(U); ref = m64.log_likelihood(beta, 1.0); m64.close()
    out = {"Q": Q}
    for on in (True, False):
        g.estep_set_tf32(on)
        mm = g.Model(ctx, X, Z, y, "binomial", "logit", precision="fp32"); mm.set_u(U)
        ctx.sync(); ctx.timer_start(); mm.set_u(U); t = ctx.timer_stop()
        out["tf32" if on else "f64_narrowed"] = {"rel_err": abs(mm.log_likelihood(beta, 1.0) - ref) / abs(ref), "set_u_ms": t}
        mm.close()
    g.estep_set_tf32(True)
    print(json.dumps(out), flush=True)

# throughput of the dense contraction zd = Z u alone (device-resident operands): fp64 DMMA (TMA kernel), fp32 mode on tcgen05 3xTF32, fp32 mode on the fp64 product
n, Q, m = 8192, 4096, 16384
X = np.asfortranarray(np.ones((n, 1))); Z = np.asfortranarray(rng.standard_normal((n, Q)) / np.sqrt(Q))
U = np.asfortranarray(rng.standard_normal((Q, m))); y = (rng.random(n) < 0.5).astype(float)
res = {"shape": [n, Q, m], "flop": 2.0 * n * Q * m}
for name, prec, tf in (("f64_dmma", "fp64", True), ("f32_tcgen05_3xtf32", "fp32", True), ("f32_via_f64_dmma", "fp32", False)):
    g.estep_set_tf32(tf)
    mm = g.Model(ctx, X, Z, y, "gaussian", "identity", precision=prec); mm.set_u(U); mm.rebuild_zd()
    ts = []
    for _ in range(3):
        ctx.sync(); ctx.timer_start(); mm.rebuild_zd(); ts.append(ctx.timer_stop())
    t = float(np.median(ts))
    res[name] = {"ms": t, "tflops_equiv": 2.0 * n * Q * m / t / 1e9}
    mm.close()
g.estep_set_tf32(True)
print(json.dumps(res), flush=True)
