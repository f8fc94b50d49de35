#!/usr/bin/env python
"""Leapfrog throughput of the structure-aware sampler (hmc_sparse.cu) against the dense kernels on the cluster designs C1, C2, C4.
One JSON line per (config, kernel):  python tools/hmc_sparse_bench.py [quick]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
ctx = g.Context(0)
NAMES = {1: "two-GEMM", 2: "on-chip dense", 3: "structure-aware"}


def run(tag, cfg, variants, n_chains, warmup, ns, lam=5.0, max_steps=100, reps=2):
    mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], cfg["family"], cfg["link"])
    for variant in variants:
        g.hmc_set_variant(variant)
        try:
            best = None
            for rep in range(reps):
                out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=warmup, nsamp_per_chain=ns, lam=lam, max_steps=max_steps,
                                     target_accept=0.95, n_chains=n_chains, seed=11, keep_on_device=True, want_u=False)
                st = out["stats"]
                if best is None or st["kernel_ms"] < best["kernel_ms"]:
                    best = st
            st = best
            print(json.dumps({"config": tag, "kernel": NAMES[st["kernel_variant"]], "n": cfg["n"], "Q": cfg["Q"], "rows_used": st["rows_used"],
                              "zl_nonzeros": st["zl_nonzeros"], "chains": n_chains, "proposals": warmup + ns, "ms": round(st["kernel_ms"], 3),
                              "leapfrog_per_s": st["leapfrog_total"] / (st["kernel_ms"] * 1e-3),
                              "ns_per_leapfrog_per_chain_slot": st["kernel_ms"] * 1e6 / (st["leapfrog_total"] / n_chains),
                              "accept": st["accept_rate"], "steps_mean": st["steps_mean"]}), flush=True)
        except g.GmbError as e:
            print(json.dumps({"config": tag, "variant": variant, "skipped": str(e)[:120]}), flush=True)
        finally:
            g.hmc_set_variant(0)
    mdl.close()


run("C2 (bench draw: 1000 chains, warm-up 500 + 10)", synth.config2(m=4), (3, 2), 1000, 500, 9)
run("C2 4000 chains", synth.config2(m=4), (3,), 4000, 100, 4)
run("C1", synth.config1(m=4), (3, 2), 1000, 100 if quick else 500, 9)
c4 = synth.config4(ncl=1000, nt=10, k=1, m=4)
run("C4 n=Q=10^4", c4, (3,), 592, 20 if quick else 100, 4)
run("C4 n=Q=10^4", c4, (1,), 592, 2, 1, reps=1)
c4s = synth.config4(ncl=100, nt=10, k=2, m=4)
run("C4 n=2000 Q=1000", c4s, (3, 1), 592, 20, 4, reps=1)
