import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth
ctx = g.Context(0)
cfg = synth.config2(m=4)
mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
for nch in (100, 125, 200, 250, 400, 500, 1000, 2000):
    per = 10000 // nch
    best = None
    for rep in range(2):
        out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=500, nsamp_per_chain=per - 1, lam=5.0, max_steps=100, target_accept=0.95,
                             n_chains=nch, seed=11 + rep, keep_on_device=True, want_u=False)
        st = out["stats"]
        if best is None or st["kernel_ms"] < best["kernel_ms"]: best = st
    print(json.dumps({"chains": nch, "cols_per_chain": per, "ms": round(best["kernel_ms"], 3), "leapfrog_per_s": best["leapfrog_total"] / best["kernel_ms"] * 1e3,
                      "ns_per_step": best["kernel_ms"] * 1e6 / (best["leapfrog_total"] / nch)}), flush=True)
