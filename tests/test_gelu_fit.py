"""The fp16 epilogues' GELU (csrc/common.cuh:gelu_fast) restated in numpy float32 and held against the exact erf GELU
(nn.GELU() default, attentions.py:239-241): the coefficients in the .cuh file are parsed, not duplicated."""
import os
import re

import numpy as np
from scipy.special import erf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coeffs():
    src = open(os.path.join(ROOT, "s2m2_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float gelu_fast(float x)"):]
    body = body[:body.index("return")]
    c = [float(v) for v in re.findall(r"(-?\d\.\d+e[+-]\d+)f", body)]
    assert len(c) == 5, c
    return np.array(c, dtype=np.float32)     # c4, c3, c2, c1, c0 (Horner order)


def gelu_fast_np(x):
    c = _coeffs()
    x = x.astype(np.float32)
    t = x * x
    p = np.full_like(x, c[0]) * t + c[1]
    for k in (2, 3, 4):
        p = (p * t + c[k]).astype(np.float32)
    with np.errstate(over="ignore"):
        e = np.exp2((x * p).astype(np.float32)).astype(np.float32)
    return (x / (np.float32(1) + e)).astype(np.float32)


def test_gelu_fast_accuracy():
    x = np.linspace(-12, 12, 1_200_001).astype(np.float32)
    ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    err = np.abs(gelu_fast_np(x) - ref)
    assert err.max() < 4e-6, err.max()
    pos = x > 0.01
    assert (err[pos] / ref[pos]).max() < 2e-5                  # 20x below fp16's 2^-11 relative rounding


def test_gelu_fast_extremes_are_finite_and_saturate():
    x = np.array([-65504.0, -1e4, -40.0, -0.0, 0.0, 40.0, 1e4, 65504.0, 3e38, -3e38], dtype=np.float32)
    y = gelu_fast_np(x)
    assert np.isfinite(y).all()
    assert (y[:4] == 0).all() and y[4] == 0
    np.testing.assert_array_equal(y[5:9], x[5:9])
