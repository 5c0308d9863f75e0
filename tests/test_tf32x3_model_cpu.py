"""Numeric model of the 3xTF32 operand split used by b200_linear_tf32x3 (csrc/mlp_tc.cu): x = hi + lo
with hi = the tf32 truncation of x and lo = x - hi (itself truncated to tf32 by the tensor core),
product ~ hi*hi' + lo*hi' + hi*lo'.  Checks in exact float64 arithmetic that the representation +
dropped-term error is below 3 * 2^-20 * sum|x w| for every input (truncation keeps 10 explicit
mantissa bits: |lo| < 2^-10 |x|, so lo*lo' < 2^-20 |x w| and the two truncated cross terms add 2^-20 each) (the measured kernel error, which also
contains the fp32 accumulation, is <= 1e-6 * sum|x w|: tests/test_gpu_linear_tc.py)."""
import numpy as np
import pytest


def _tf32_trunc(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return b.view(np.float32)


@pytest.mark.parametrize("din", [32, 1792])
def test_three_product_split_error_bound(din):
    rng = np.random.default_rng(din)
    x = (rng.standard_normal((64, din)) * np.exp(rng.uniform(-6, 6, (64, din)))).astype(np.float32)   # wide dynamic range
    w = (rng.standard_normal((din, 16)) * np.exp(rng.uniform(-6, 6, (din, 16)))).astype(np.float32)
    xh, wh = _tf32_trunc(x), _tf32_trunc(w)
    xl, wl = _tf32_trunc(x - xh), _tf32_trunc(w - wh)        # what the tensor core sees of the lo parts
    assert (np.abs(x - xh) <= np.abs(x) * 2.0 ** -10).all()  # 10 explicit mantissa bits kept
    f = np.float64
    approx = xh.astype(f) @ wh.astype(f) + xl.astype(f) @ wh.astype(f) + xh.astype(f) @ wl.astype(f)
    exact = x.astype(f) @ w.astype(f)
    mag = np.abs(x).astype(f) @ np.abs(w).astype(f)
    err = np.abs(approx - exact) / mag
    assert err.max() <= 3 * 2.0 ** -20, float(err.max())
    # a single tf32 product (no split) is three orders of magnitude worse: why the split exists
    single = np.abs(xh.astype(f) @ wh.astype(f) - exact) / mag
    assert single.max() > 50 * err.max()
