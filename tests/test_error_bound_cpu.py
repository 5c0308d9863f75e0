"""The fused scorer's exactness argument rests on |coarse - exact| <= eps with
eps = ERR_COEF * ||u|| * max_i ||i|| (csrc/score_topk_tc.cu): coarse = dot product of the bf16-rounded
vectors accumulated in fp32 on the tensor core, exact = the fp32 fma chain.  This test checks the
coefficient on the CPU with an exact (float64) model of both sides: random vectors, adversarial
vectors whose every component sits at the worst rounding position, widths up to the kernel's limit,
and a pessimistic model of the accumulator (truncation after every one of the d products)."""
import re

import numpy as np
import pytest


def _err_coef():
    src = open("librecommender_b200/csrc/score_topk_tc.cu").read()
    return float(re.search(r"ERR_COEF\s*=\s*([0-9.eE+-]+)f", src).group(1))


def _bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32 (numpy has no bf16)."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000
    return b.astype(np.uint32).view(np.float32)


def _trunc_accumulate(prods):
    """fp32 accumulator that TRUNCATES (toward zero) after every addition: the most pessimistic
    reading of the tensor core's accumulation (it adds several products per step in a wider adder)."""
    acc = np.float64(0.0)
    for p in prods:
        acc = acc + p
        m, e = np.frexp(acc)
        acc = np.ldexp(np.trunc(m * 2.0 ** 24) / 2.0 ** 24, e)
    return acc


@pytest.mark.parametrize("d", [7, 64, 128, 256])
def test_bf16_coarse_score_error_is_below_eps(d):
    coef = _err_coef()
    assert coef >= 2.0 ** -7 * (1 + 2.0 ** -9)
    rng = np.random.default_rng(d)
    worst = 0.0
    cases = []
    for _ in range(300):
        cases.append((rng.standard_normal(d), rng.standard_normal(d)))
    # adversarial: every component exactly half-way between two bf16 values (max rounding error),
    # all errors with the same sign
    base = 1.0 + 2.0 ** -8                      # halfway between 1 and 1 + 2^-7 in bf16
    for s in (1.0, 0.37, 11.0):
        cases.append((np.full(d, base * s), np.full(d, base / s)))
        cases.append((np.full(d, base * s) * (-1) ** np.arange(d), np.full(d, base / s) * (-1) ** np.arange(d)))
    for u, i in cases:
        u32, i32 = u.astype(np.float32), i.astype(np.float32)
        exact = np.float64(0.0)
        for k in range(d):                        # the kernel's exact-score definition (fp32 fma chain)
            exact = np.float64(np.float32(np.float64(u32[k]) * np.float64(i32[k]) + exact))
        prods = _bf16(u32).astype(np.float64) * _bf16(i32).astype(np.float64)
        coarse = _trunc_accumulate(prods)
        bound = coef * float(np.linalg.norm(u32.astype(np.float64))) * float(np.linalg.norm(i32.astype(np.float64)))
        worst = max(worst, abs(coarse - exact) / bound)
        assert abs(coarse - exact) <= bound, (d, abs(coarse - exact), bound)
    assert worst < 1.0
