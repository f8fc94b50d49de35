#!/usr/bin/env python
"""One-off check (CPU, ~10 min): the oracle loop (oracle/mcml_loop.py) against tests/golden/REFSRC_mcml_full.npz — the reference's own
src/mcml_full.cpp — for ALL seeds and both configurations of tests/test_gpu_fit_parity.py.  Prints the worst differences; the GPU test's
direct comparison with the golden file (library vs reference source) is implied by its comparison with the oracle loop plus these."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from glmmrmcml_b200 import synth          # noqa: E402
from oracle import mcml_loop               # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_refsrc import CASES, SEEDS   # noqa: E402

gold = np.load(os.path.join(ROOT, "tests", "golden", "REFSRC_mcml_full.npz"))
for name, (make, tail, kw) in CASES.items():
    cfg = make()
    start = np.concatenate([cfg["beta"] * 0.8, tail])
    a = (cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start)
    wb = wt = wu = 0.0
    for k, seed in enumerate(SEEDS):
        o = mcml_loop.mcml_full(*a, seed=seed, **kw)
        db = float(np.max(np.abs(o["beta"] - gold[name + "_beta"][k]))); dt = float(np.max(np.abs(o["theta"] - gold[name + "_theta"][k])))
        du = float(np.max(np.abs(o["u"][:, -1] - gold[name + "_u_last_column"][k])))
        assert bool(gold[name + "_converged"][k]) == o["converged"], (name, seed)
        wb, wt, wu = max(wb, db), max(wt, dt), max(wu, du)
        print(name, seed, "beta %.2e theta %.2e u %.2e" % (db, dt, du), flush=True)
    print(name, "WORST beta %.2e theta %.2e u(last column) %.2e" % (wb, wt, wu), flush=True)
