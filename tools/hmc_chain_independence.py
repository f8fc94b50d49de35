#!/usr/bin/env python
"""Diagnostic: are the chains of one sampling run independent?  Ratio of the between-run sd of posterior means (6 seeds) to the
standard error predicted from per-chain means, and the correlation of per-chain means between neighbouring chains (both ~1 and ~0)."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import glmmrmcml_b200 as g
from glmmrmcml_b200 import synth
ctx = g.Context(0); cfg = synth.config2(m=64)
mdl = g.Model(ctx, cfg["X"], cfg["Z"], cfg["y"], "binomial", "logit")
Q = cfg["Q"]
def run(seed, variant=0, cs=0, nch=250, ns=39, warm=500):
    g.hmc_set_variant(variant); g.hmc_set_cluster_size(cs)
    out = mdl.hmc_sample(cfg["L"], cfg["beta"], 1.0, warmup=warm, nsamp_per_chain=ns, lam=5.0, max_steps=100, target_accept=0.95,
                         n_chains=nch, seed=seed, want_u=True)
    Uc = out["u"].reshape(Q, ns + 1, nch, order="F").transpose(0, 2, 1)[:, :, 1:]
    cm = Uc.mean(axis=2)
    return cm.mean(axis=1), cm.std(axis=1, ddof=1) / np.sqrt(nch), out["stats"], cm
for label, kw in (("auto (structure-aware)", dict()), ("on-chip dense", dict(variant=2)), ("on-chip dense cs1", dict(variant=2, cs=1)), ("two-gemm", dict(variant=1, nch=64, ns=20, warm=200))):
    ms, ss = [], []
    for seed in range(1, 7):
        m_, s_, st, cm = run(seed, **kw)
        ms.append(m_); ss.append(s_)
    ms = np.array(ms); ss = np.array(ss)
    between = ms.std(axis=0, ddof=1)
    pred = np.sqrt((ss ** 2).mean(axis=0))
    print(label, "accept", st["accept_rate"], "ratio between-run sd / predicted se: median %.2f max %.2f" % (np.median(between / pred), np.max(between / pred)))
    # correlation between chains inside one run: adjacent chains in a group
    c = np.corrcoef(cm[:, 0::2][:, :100].ravel(), cm[:, 1::2][:, :100].ravel())[0, 1]
    print("   corr of per-chain means between neighbouring chains (all coords pooled, centred per coord):",
          np.corrcoef((cm - cm.mean(axis=1, keepdims=True))[:, 0::2][:, :100].ravel(), (cm - cm.mean(axis=1, keepdims=True))[:, 1::2][:, :100].ravel())[0, 1])
