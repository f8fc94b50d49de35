"""The C-ABI library without a GPU: it loads, exports every symbol include/glmmrmcml_b200.h declares, fails loudly when
no CUDA device exists, and its host-only entry points (covariance shape, optimiser, finite-difference stencils) work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import glmmrmcml_b200 as g
from glmmrmcml_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "glmmrmcml_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gmb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = g.lib()
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/glmmrmcml_b200.h but not exported"
    # and the Python binding table covers the header
    assert set(syms) <= set(_lib.PROTOTYPES), sorted(set(syms) - set(_lib.PROTOTYPES))
    assert "sm_100a" in g.version()


def test_no_cpu_fallback():
    """Without a CUDA device every compute path must fail with GMB_ECUDA (code 3), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g.GmbError) as e:
        g.Context(0)
    assert e.value.code == _lib.GMB_ECUDA and "no CPU fallback" in str(e.value)
    cfg = synth.config2(m=4)
    with pytest.raises(g.GmbError) as e:
        g.mvn_ll(cfg["cov"], cfg["data"], cfg["eff_range"], cfg["theta"], cfg["U"])
    assert e.value.code == _lib.GMB_ECUDA


def test_product_code_does_not_touch_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "glmmrmcml_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "libref" not in txt, f


def test_cov_shape():
    for make, want in ((lambda: synth.config1(m=2), (60, 60, 2)), (lambda: synth.config2(m=2), (10, 50, 2)),
                       (lambda: synth.config3(nloc=17, m=2), (1, 17, 2)), (lambda: synth.config4(ncl=7, nt=3, m=2), (7, 21, 2))):
        assert g.cov_shape(make()["cov"]) == want
    with pytest.raises(g.GmbError):
        g.cov_shape(np.array([[0, 3, 99, 1, 0]], dtype=np.int32))


def test_minimize_bounded_matches_scipy():
    from scipy.optimize import minimize
    rng = np.random.default_rng(0)
    for n in (2, 5, 9):
        A = rng.standard_normal((n, n)); A = A @ A.T + n * np.eye(n); b = rng.standard_normal(n)
        f = lambda X: 0.5 * np.einsum("ik,ij,jk->k", X, A, X) - b @ X + 0.05 * np.sum(X ** 4, axis=0)
        x0 = rng.standard_normal(n) + 2
        r = g.minimize_bounded(f, x0)
        ref = minimize(lambda x: f(x[:, None])[0], x0, method="BFGS", options=dict(gtol=1e-10))
        assert np.max(np.abs(r["x"] - ref.x)) < 5e-6
        lo = np.full(n, -np.inf); lo[0] = ref.x[0] + 0.3                       # make one bound active
        r = g.minimize_bounded(f, x0, lower=lo)
        refb = minimize(lambda x: f(x[:, None])[0], np.maximum(x0, lo), method="L-BFGS-B", bounds=[(l if np.isfinite(l) else None, None) for l in lo],
                        options=dict(ftol=1e-15, gtol=1e-10))
        assert abs(r["x"][0] - lo[0]) < 1e-12 and np.max(np.abs(r["x"] - refb.x)) < 5e-5


def test_minimize_handles_infinite_objective_regions():
    # like -mvn_ll(theta): +inf where D(theta) is not positive definite
    f = lambda X: np.where(X[0] < 1.0, (X[0] - 0.8) ** 2 + (X[1] - 0.3) ** 2, np.inf)
    r = g.minimize_bounded(f, [0.25, 0.9], lower=[1e-6, 1e-6])
    assert np.max(np.abs(r["x"] - [0.8, 0.3])) < 1e-5


def test_fd_hessian_is_the_optimhess_stencil():
    rng = np.random.default_rng(3)
    n = 4
    A = rng.standard_normal((n, n)); A = A @ A.T + np.eye(n)
    f = lambda X: 0.5 * np.einsum("ik,ij,jk->k", X, A, X) + np.sum(np.sin(X), axis=0)
    x = rng.standard_normal(n)
    H, nfev = g.fd_hessian(f, x, 1e-4)
    assert nfev == 4 * n * n
    assert np.allclose(H, A - np.diag(np.sin(x)), rtol=1e-5, atol=1e-6)
    assert np.array_equal(H, H.T)
    # python transcription of R's optimhess (optim.c) for the same function
    def grad(p):
        out = np.zeros(n)
        for i in range(n):
            e = np.zeros(n); e[i] = 1e-4
            out[i] = (f((p + e)[:, None])[0] - f((p - e)[:, None])[0]) / 2e-4
        return out
    Hr = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = 1e-4
        Hr[i] = (grad(x + e) - grad(x - e)) / 2e-4
    Hr = 0.5 * (Hr + Hr.T)
    assert np.allclose(H, Hr, rtol=1e-9, atol=1e-10)


def test_objective_error_aborts_cleanly():
    def bad(X):
        raise RuntimeError("boom")
    with pytest.raises(g.GmbError):
        g.minimize_bounded(bad, [1.0, 2.0])


C_CONSUMER = r"""
/* a C99 consumer of the boundary, as cgo / a .Call shim / any FFI generator would see it */
#include <stdio.h>
#include <string.h>
#include "glmmrmcml_b200.h"
int main(void) {
    int B = 0, Q = 0, R = 0;
    int32_t cov[10] = {0, 1, 3, 3, 1, 1, 1, 1, 0, 1};   /* column-major 2 x 5: blocks 0 and 1 of dimension 3, each gr (id 1) on 1 variable, parameters theta[0] and theta[1] */
    gmb_ctx* ctx = NULL;
    if (strstr(gmb_version(), "sm_100a") == NULL) return 10;
    if (gmb_cov_shape(cov, 2, &B, &Q, &R) != GMB_OK) { fprintf(stderr, "%s\n", gmb_last_error()); return 11; }
    printf("B=%d Q=%d R=%d\n", B, Q, R);
    cov[4] = 99;                                        /* unknown covariance function id */
    if (gmb_cov_shape(cov, 2, &B, &Q, &R) == GMB_OK || strlen(gmb_last_error()) == 0) return 12;
    if (gmb_ctx_create(0, &ctx) == GMB_OK) { gmb_ctx_destroy(ctx); printf("device present\n"); }
    else printf("no device: %s\n", gmb_last_error());
    return 0;
}
"""


def test_header_is_plain_c_and_links_from_a_c_program(tmp_path):
    """include/glmmrmcml_b200.h compiled by a C compiler (-std=c99 -pedantic -Werror: no C++ in the boundary), linked against the
    shared library, host-only entry points called from C."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "consumer.c"
    src.write_text(C_CONSUMER)
    exe = tmp_path / "consumer"
    libdir = os.path.join(ROOT, "glmmrmcml_b200")
    r = subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lglmmrmcml_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "B=2 Q=6 R=2" in r.stdout


def test_missing_library_fails_loudly():
    """Without the built .so the package raises at first use — there is nothing to fall back to."""
    import subprocess
    import sys
    code = ("import numpy as np, glmmrmcml_b200 as g\n"
            "try:\n    g.mvn_ll(np.array([[0, 1, 1, 1, 0]], dtype=np.int32), np.zeros(1), np.zeros(1), np.ones(1), np.zeros((1, 2)))\n"
            "except Exception as e:\n    print('RAISED', type(e).__name__, e)\n")
    env = dict(os.environ, GMB_LIB="/nonexistent/libglmmrmcml_b200.so", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert "RAISED" in r.stdout and "libglmmrmcml_b200" in r.stdout, (r.stdout, r.stderr)
