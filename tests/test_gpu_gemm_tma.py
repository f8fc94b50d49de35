"""The TMA-staged DMMA GEMM (csrc/gemm_tma.cuh) on every contraction of the library, including shapes far smaller than its 128 x 128 x 32
tiles (edge boxes are zero-filled by the TMA unit): the parity suites of the covariance path (Cholesky trailing updates, blocked forward
substitution), the E-step (zd = Z u with a dense Z) and the two-contraction sampler (fused-epilogue products, triangular k-tile skipping) are
re-run in a child process with GMB_GEMM_TMA=2, which sends every product through the TMA kernel instead of only the large ones."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_suites_with_every_product_on_the_tma_kernel():
    env = dict(os.environ, GMB_GEMM_TMA="2")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_cov.py", "tests/test_gpu_estep.py", "tests/test_gpu_hmc.py",
                        "tests/test_gpu_families.py", "tests/test_laplace.py"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
