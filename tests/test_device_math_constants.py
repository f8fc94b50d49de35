"""The table-driven exp of the device code, checked on the CPU: the constants are PARSED from the CUDA sources (exp_table.cuh, common.cuh
`dev_exp_tab`, hmc_sparse_common.cuh `SP_EXPC16`) and the two algorithms are replayed in IEEE double arithmetic with an exact fused
multiply-add (rational arithmetic, rounded once), against mpmath.  Pins the accuracy statement of DESIGN.md §3 (K2/K3 arithmetic, K6s: error
below 1.5 ulp of the exact value — measured maxima 1.2 ulp for the 64-entry / degree-5 form, 1.02 ulp for the 16-entry / degree-7 form; the
library exp is at 0.5) without a GPU; the GPU tests then show that the kernels using these functions agree with the oracle."""
import math
import os
import re
from fractions import Fraction

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
mp.mp.prec = 200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "glmmrmcml_b200", "csrc")


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))      # exact product and sum, one rounding (Fraction -> float rounds to nearest even)


def table64():
    txt = open(os.path.join(CSRC, "exp_table.cuh")).read()
    body = txt[txt.index("{") + 1: txt.index("}")]
    vals = [float(v) for v in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", body)]
    assert len(vals) == 64
    return vals


def ulp_err(got, x):
    want = mp.exp(mp.mpf(x))
    u = math.ulp(float(want))
    return abs(mp.mpf(got) - want) / u


def test_table_entries_are_correctly_rounded():
    for j, v in enumerate(table64()):
        assert v == float(mp.power(2, mp.mpf(j) / 64)), j


def _scale(m, k, shift):
    """__hiloint2double(__double2hiint(m) + ((k >> shift) << 20), lo): adds k >> shift to the binary exponent of m in [1, 2)."""
    return math.ldexp(m, k >> shift)


def exp_tab64(x, T, c_inv, c_hi, c_lo):
    t = fma(x, c_inv, 6755399441055744.0)
    k = int(np.float64(t).view(np.int64) & 0xFFFFFFFF)
    k = k - (1 << 32) if k >= (1 << 31) else k                 # __double2loint: the low word as a signed int
    t -= 6755399441055744.0
    r = fma(t, c_hi, x)
    r = fma(t, c_lo, r)
    Tj = T[k & 63]
    q = fma(r, 1.0 / 120.0, 1.0 / 24.0)
    q = fma(q, r, 1.0 / 6.0)
    q = fma(q, r, 0.5)
    q = fma(q, r, 1.0)
    q = q * r
    m = fma(Tj, q, Tj)
    k = min(max(k, -64512), 64512)
    return _scale(m, k, 6)


def exp_tab16(x, T16, cc):
    t = fma(x, cc[0], cc[1])
    k = int(np.float64(t).view(np.int64) & 0xFFFFFFFF)
    k = k - (1 << 32) if k >= (1 << 31) else k
    t -= cc[1]
    Tj = T16[k & 15]
    r = fma(t, cc[2], x)
    r = fma(t, cc[3], r)
    r2 = r * r
    a = fma(r, cc[4], cc[5]); b = fma(r, cc[6], cc[7]); d3 = fma(r, cc[8], cc[9])
    q = fma(a, r2, b)
    q = fma(q, r2, d3)
    q = fma(q, r, cc[10])
    q = q * r
    m = fma(Tj, q, Tj)
    k = min(max(k, -16128), 16128)
    return _scale(m, k, 4)


def _consts64():
    txt = open(os.path.join(CSRC, "common.cuh")).read()
    body = txt[txt.index("__device__ __forceinline__ double dev_exp_tab("):]
    body = body[: body.index("\n}\n")]
    c_inv = float(re.search(r"fma\(x, ([-0-9.eE+]+), 6755399441055744\.0\)", body).group(1))
    c_hi = float(re.search(r"fma\(t, ([-0-9.eE+]+), x\)", body).group(1))
    c_lo = float(re.search(r"fma\(t, ([-0-9.eE+]+), r\)", body).group(1))
    assert "min(max(k, -64512), 64512)" in body and "(k >> 6) << 20" in body
    return c_inv, c_hi, c_lo


def _consts16():
    txt = open(os.path.join(CSRC, "hmc_sparse_common.cuh")).read()
    body = txt[txt.index("SP_EXPC16[12] = {") + len("SP_EXPC16[12] = {"):]
    body = re.sub(r"//.*", "", body[: body.index("};")])
    vals = [eval(v, {"__builtins__": {}}) for v in body.replace("\n", " ").split(",") if v.strip()]      # entries like 1.0 / 5040.0
    assert len(vals) == 12
    return [float(v) for v in vals]


def _sample_points(rng, n):
    xs = np.concatenate([rng.uniform(-690, 690, n), rng.uniform(-40, 40, n), rng.normal(0, 1e-3, n // 4),
                         np.arange(-64, 65) * (math.log(2) / 64), np.arange(-64, 65) * (math.log(2) / 128)])   # table boundaries and mid-points
    return xs


def test_split_of_ln2_is_exact_enough():
    c_inv, c_hi, c_lo = _consts64()
    ln2_64 = mp.log(2) / 64
    assert abs(mp.mpf(c_inv) * ln2_64 - 1) < mp.mpf(2) ** -52
    assert abs(-(mp.mpf(c_hi) + mp.mpf(c_lo)) - ln2_64) / ln2_64 < mp.mpf(2) ** -75          # hi + lo carries ~ 77 bits of ln2/64
    mant = int(np.float64(abs(c_hi)).view(np.int64)) & ((1 << 52) - 1)
    assert mant % (1 << 24) == 0                                                             # 24 trailing zero bits: t * hi is exact for |t| < 2^24
    cc = _consts16()
    assert abs(-(mp.mpf(cc[2]) + mp.mpf(cc[3])) - mp.log(2) / 16) / (mp.log(2) / 16) < mp.mpf(2) ** -75
    assert (int(np.float64(abs(cc[2])).view(np.int64)) & ((1 << 52) - 1)) % (1 << 24) == 0


MAX_ULP = 1.5


def test_dev_exp_tab_accuracy():
    T = table64(); c = _consts64()
    rng = np.random.default_rng(0)
    worst = max(ulp_err(exp_tab64(float(x), T, *c), float(x)) for x in _sample_points(rng, 1500))
    assert worst <= MAX_ULP, float(worst)
    # saturation beyond e^+-698.7 (the binary exponent is clamped at +-1008): finite, monotone side
    assert exp_tab64(800.0, T, *c) >= exp_tab64(698.0, T, *c) * 0.5 and math.isfinite(exp_tab64(800.0, T, *c))
    assert 0.0 < exp_tab64(-800.0, T, *c) < 1e-300


def test_sixteen_entry_exp_accuracy():
    T16 = table64()[::4]; cc = _consts16()
    rng = np.random.default_rng(1)
    worst = max(ulp_err(exp_tab16(float(x), T16, cc), float(x)) for x in _sample_points(rng, 1500))
    assert worst <= MAX_ULP, float(worst)
