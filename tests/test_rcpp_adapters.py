"""The boundary from the reference's side: the Rcpp adapters of src/ (mcml_full.cpp, mcml_optim.cpp, mcml_la.cpp — the reference's exported
signatures, bodies forwarding to the C-ABI), compiled against the stand-in Rcpp / RcppEigen headers of oracle/shim and driven by
tests/adapters_driver.cpp the way src/RcppExports.cpp drives them.  CPU: they build, load, and turn the library's status into an R error.
GPU: every in-scope export returns what the ctypes mirror of the same entry point returns."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from glmmrmcml_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def drv():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "src")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    import glmmrmcml_b200 as g
    g.lib()                                                     # libglmmrmcml_b200.so first (RTLD_GLOBAL), then the adapters that link to it
    lib = C.CDLL(os.path.join(ROOT, "tests", "_build", "libgmb_rcpp_adapters.so"))
    lib.drv_last_error.restype = C.c_char_p
    lib.drv_set_seed.argtypes = [C.c_ulonglong]
    return lib


def _cov(cfg):
    cov = np.asfortranarray(np.asarray(cfg["cov"], dtype=np.int32).reshape(-1, 5))
    data = np.ascontiguousarray(cfg["data"], dtype=np.float64)
    eff = np.ascontiguousarray(cfg["eff_range"], dtype=np.float64)
    return (cov, data, eff), [cov.ctypes.data_as(ip), C.c_int(cov.shape[0]), _d(data), C.c_int(data.size), _d(eff), C.c_int(eff.size)]


def _model(cfg):
    Z = np.asfortranarray(cfg["Z"]); X = np.asfortranarray(cfg["X"]); y = np.ascontiguousarray(cfg["y"]); U = np.asfortranarray(cfg["U"])
    return (Z, X, y, U), [_d(Z), _d(X), _d(y)], _d(U)


def test_adapters_build_and_raise_r_errors_without_a_gpu(drv):
    """Rcpp::stop(gmb_last_error()) reaches the caller as an R error: without a CUDA device, the 'no CPU fallback' message; with one,
    an unknown family/link (the reference throws std::out_of_range from string_to_case.at, mcmlmodel.h:89)."""
    import torch
    cfg = synth.config2(m=6, ncl=4, nt=3, nind=2)
    keep, cargs = _cov(cfg)
    (Z, X, y, U), margs, up = _model(cfg)
    out = C.c_double()
    if not torch.cuda.is_available():
        rc = drv.drv_mvn_ll(*cargs, _d(cfg["theta"]), C.c_int(2), up, C.c_int(cfg["Q"]), C.c_int(6), C.byref(out))
        assert rc == 1 and b"no CPU fallback" in drv.drv_last_error()
    else:
        beta = np.zeros(cfg["P"]); theta = np.zeros(2); sg = C.c_double()
        start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
        rc = drv.drv_mcml_optim(*cargs, *margs, up, C.c_int(cfg["n"]), C.c_int(cfg["P"]), C.c_int(cfg["Q"]), C.c_int(6), b"binomial", b"cloglog",
                                _d(start), C.c_int(start.size), C.c_int(0), C.c_int(0), C.c_int(0), _d(beta), _d(theta), C.byref(sg))
        assert rc == 1 and b"unknown family/link" in drv.drv_last_error()


@pytest.mark.gpu
def test_every_export_matches_the_ctypes_mirror(drv, gctx):
    import glmmrmcml_b200 as g
    gctx.make_default()
    cfg = synth.config2(m=300, seed=11, ncl=8, nt=4, nind=6)
    n, P, Q, m = cfg["n"], cfg["P"], cfg["Q"], 300
    keep, cargs = _cov(cfg)
    (Z, X, y, U), margs, up = _model(cfg)
    dims = [C.c_int(n), C.c_int(P), C.c_int(Q), C.c_int(m)]
    fam, link = b"binomial", b"logit"
    start = np.concatenate([cfg["beta"], cfg["theta"], [1.0]])
    a = (cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], cfg["U"], "binomial", "logit")
    # mvn_ll
    out = C.c_double()
    assert drv.drv_mvn_ll(*cargs, _d(cfg["theta"]), C.c_int(2), up, C.c_int(Q), C.c_int(m), C.byref(out)) == 0, drv.drv_last_error()
    assert out.value == g.mvn_ll(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], cfg["U"])
    # mcml_optim (MCNR and MCEM), mcml_simlik
    for mcnr, simlik in ((1, 0), (0, 0), (0, 1)):
        beta = np.zeros(P); theta = np.zeros(2); sg = C.c_double()
        rc = drv.drv_mcml_optim(*cargs, *margs, up, *dims, fam, link, _d(start), C.c_int(start.size), C.c_int(0), C.c_int(mcnr), C.c_int(simlik),
                                _d(beta), _d(theta), C.byref(sg))
        assert rc == 0, drv.drv_last_error()
        ref = g.mcml_simlik(*a, start) if simlik else g.mcml_optim(*a, start, 0, bool(mcnr))
        assert np.array_equal(beta, ref["beta"]) and np.array_equal(theta, ref["theta"]) and sg.value == ref["sigma"]
    # mcml_hess, aic_mcml
    k = P + 2
    H = np.zeros((k, k), order="F")
    assert drv.drv_mcml_hess(*cargs, *margs, up, *dims, fam, link, _d(start), C.c_int(k), C.c_double(1e-4), C.c_int(0), _d(H), C.c_int(k)) == 0, drv.drv_last_error()
    assert np.array_equal(H, g.mcml_hess(*a, start[:k], 1e-4, 0))
    assert drv.drv_aic_mcml(*cargs, *margs, up, *dims, fam, link, _d(cfg["beta"]), C.c_int(P), _d(cfg["theta"]), C.c_int(2), C.byref(out)) == 0
    assert out.value == g.aic_mcml(*a, cfg["beta"], cfg["theta"])
    # mcmc_sample: the adapter draws its seed from R's generator (unif_rand under RNGScope): same R seed -> same samples
    L = np.asfortranarray(cfg["L"])
    S = [np.zeros((Q, 65), order="F") for _ in range(3)]
    for i, seed in enumerate((7, 7, 8)):
        drv.drv_set_seed(seed)
        rc = drv.drv_mcmc_sample(margs[0], _d(L), margs[1], margs[2], _d(cfg["beta"]), C.c_int(n), C.c_int(P), C.c_int(Q), fam, link, C.c_int(50), C.c_int(64),
                                 C.c_double(1.0), C.c_double(1.0), C.c_int(0), C.c_int(500), C.c_int(20), C.c_double(0.9), _d(S[i]))
        assert rc == 0, drv.drv_last_error()
    assert np.array_equal(S[0], S[1]) and not np.array_equal(S[0], S[2]) and np.all(np.isfinite(S[0])) and np.std(S[0]) > 0
    # mcml_full
    beta = np.zeros(P); theta = np.zeros(2); sg = C.c_double(); conv = C.c_int(); u = np.zeros((Q, 401), order="F")
    drv.drv_set_seed(3)
    rc = drv.drv_mcml_full(*cargs, *margs, C.c_int(n), C.c_int(P), C.c_int(Q), fam, link, _d(start), C.c_int(start.size), C.c_int(1), C.c_int(400), C.c_int(4),
                           C.c_int(60), C.c_double(1e-2), C.c_double(1.0), C.c_int(20), C.c_double(0.9), _d(beta), _d(theta), C.byref(sg), C.byref(conv), _d(u))
    assert rc == 0, drv.drv_last_error()
    assert np.all(np.isfinite(beta)) and np.all(theta > 0) and np.all(np.isfinite(u)) and np.std(u) > 0
    # mcml_la / mcml_la_nr
    for nr in (0, 1):
        beta = np.zeros(P); theta = np.zeros(2); sg = C.c_double(); se = np.zeros(start.size); uu = np.zeros(Q)
        rc = drv.drv_mcml_la(*cargs, *margs, C.c_int(n), C.c_int(P), C.c_int(Q), fam, link, _d(start), C.c_int(start.size), C.c_int(nr), C.c_int(0), C.c_double(1e-3),
                             C.c_int(3), _d(beta), _d(theta), C.byref(sg), _d(se), _d(uu))
        assert rc == 0, drv.drv_last_error()
        ref = (g.mcml_la_nr if nr else g.mcml_la)(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["Z"], cfg["X"], cfg["y"], "binomial", "logit", start,
                                                  usehess=False, tol=1e-3, verbose=False, maxiter=3)
        assert np.array_equal(beta, ref["beta"]) and np.array_equal(theta, ref["theta"]) and np.array_equal(uu, ref["u"].ravel())
