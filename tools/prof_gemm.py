"""Profiling driver: zd = Z u as a dense n x Q x m contraction (sparse-Z gather off), for ncu captures of the DMMA GEMM kernels."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glmmrmcml_b200 as g
n, Q, m = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 4096, 8192)
rng = np.random.default_rng(1)
X = np.ones((n, 1)); Z = np.asfortranarray(rng.standard_normal((n, Q))); y = rng.standard_normal(n)
U = np.asfortranarray(rng.standard_normal((Q, m)))
ctx = g.Context(0)
g.estep_set_sparse_zd(False)
mdl = g.Model(ctx, X, Z, y, "gaussian", "identity")
mdl.set_u(U)
mdl.set_u(U)
# device-only timing: the GEMM again through use_device_u after invalidating (hmc not needed): time set_u minus copy is not separable, so time 3 set_u and report
import time
ctx.sync(); t0 = time.perf_counter(); mdl.set_u(U); ctx.sync(); t1 = time.perf_counter()
print("set_u wall ms", (t1 - t0) * 1e3, "GEMM flops", 2.0 * n * Q * m)
zd = None
