"""The GELU (erf form) every fast kernel evaluates: 0.5 x + |x| (0.5 - 2^-Q(|x|)) with the degree-5 polynomial Q whose
coefficients live in csrc/common.cuh (DWB_GELU_Q0..Q5, stored negated).  CPU check of those constants against
torch.nn.functional.gelu's definition (models/sashimi.py:66, s4.py:1419 use nn.GELU() = erf form) in the same fp32 arithmetic."""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coefficients():
    src = open(os.path.join(ROOT, "diffwave_sashimi_b200", "csrc", "common.cuh")).read()
    q = [float(re.search(r"#define DWB_GELU_Q%d\s+(-?[0-9.eE+-]+)f" % i, src).group(1)) for i in range(6)]
    return [np.float32(v) for v in q]


def _gelu_fast(x, q):
    x = x.astype(np.float32)
    a = np.abs(x)
    p = q[5] * a + q[4]
    for i in (3, 2, 1, 0):
        p = p * a + q[i]
    h = np.exp2(p).astype(np.float32)                 # Phi(-|x|)
    return a * (np.float32(0.5) - h) + np.float32(0.5) * x


def _gelu_exact(x):
    return np.array([0.5 * v * math.erfc(-v / math.sqrt(2.0)) for v in x])


def test_gelu_polynomial_matches_erf_form():
    q = _coefficients()
    x = np.linspace(-12.0, 12.0, 240001)
    err = np.abs(_gelu_fast(x, q).astype(np.float64) - _gelu_exact(x))
    assert err.max() < 1.2e-6, err.max()
    assert np.sqrt((err[np.abs(x) < 4] ** 2).mean()) < 4e-7


def test_gelu_polynomial_is_safe_far_out():
    """-Q must keep decreasing on the whole half line (no clamp in the kernels): huge |x| underflows 2^-Q to 0."""
    q = [float(v) for v in _coefficients()]
    a = np.logspace(-3, 7, 20001)
    p = np.polyval(q[::-1], a)
    assert np.all(np.diff(p) < 0) and p.max() < -1.0
    big = np.array([-1e30, -1e10, -100.0, -20.0, 20.0, 100.0, 1e10, 1e30])
    with np.errstate(over="ignore"):
        y = _gelu_fast(big, _coefficients())
    assert np.array_equal(y, np.maximum(big, 0).astype(np.float32))
