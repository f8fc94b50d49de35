#!/usr/bin/env python
"""Golden FITS of the reference's own src/mcml_full.cpp (oracle/_ref/librefsrc.so: the reference's source files compiled unmodified against
oracle/shim, see oracle/refsrc_driver.cpp) at exactly the settings and seeds of tests/test_gpu_fit_parity.py.  Needs /root/reference at build
time (make -C oracle ref); the output tests/golden/REFSRC_mcml_full.npz travels, so the GPU tests can compare gmb_mcml_full with the
reference's own loop without the reference tree.  Run from the repo root: python tests/golden/make_golden_refsrc.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from glmmrmcml_b200 import synth          # noqa: E402
from oracle import refsrc                  # noqa: E402

SEEDS = list(range(101, 111))
CASES = {
    # name: (config, start tail (theta, sigma), keyword arguments) — tests/test_gpu_fit_parity.py
    "C2_mcnr": (lambda: synth.config2(m=8), [0.3, 0.6, 1.0], dict(mcnr=True, m=250, maxiter=6, warmup=150, tol=1e-2, lam=5.0, maxsteps=100, target_accept=0.95)),
    "C1_mcem": (lambda: synth.config1(m=8), [0.3, 0.2, 1.0], dict(mcnr=False, m=250, maxiter=4, warmup=150, tol=5e-3, lam=5.0, maxsteps=100, target_accept=0.95)),
}


def main():
    out = {"seeds": np.array(SEEDS)}
    for name, (make, tail, kw) in CASES.items():
        cfg = make()
        start = np.concatenate([cfg["beta"] * 0.8, tail])
        a = (cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["family"], cfg["link"], start)
        B, T, S, Cv, Ul = [], [], [], [], []
        for seed in SEEDS:
            t0 = time.time()
            f = refsrc.mcml_full(*a, seed=seed, **kw)
            B.append(f["beta"]); T.append(f["theta"]); S.append(f["sigma"]); Cv.append(f["converged"]); Ul.append(f["u"][:, -1].copy())
            print(name, seed, f["beta"].round(4), f["theta"].round(4), f["converged"], "%.1f s" % (time.time() - t0), flush=True)
        out[name + "_beta"] = np.array(B); out[name + "_theta"] = np.array(T); out[name + "_sigma"] = np.array(S)
        out[name + "_converged"] = np.array(Cv); out[name + "_u_last_column"] = np.array(Ul)
        out[name + "_start"] = start
        out[name + "_y"] = cfg["y"].copy(); out[name + "_X"] = cfg["X"].copy()      # so that a comparison can check it runs on the same data
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "REFSRC_mcml_full.npz"), **out)
    print("written")


if __name__ == "__main__":
    main()
