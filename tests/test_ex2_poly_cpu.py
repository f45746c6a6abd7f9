"""CPU: float32 emulation of ptx.cuh's ex2_poly (the FMA-pipe 2^x used for a fraction of the attention exponentials when
built with -DIMAGD_ATTN_POLY=n): relative error vs exp2 over the kernel's argument range, exactness at integers, and the
clamp at the bottom."""
import numpy as np


def ex2_poly(x):
    x = np.maximum(x.astype(np.float32), np.float32(-126.0))
    magic = np.float32(12582912.0)
    r = (x + magic).astype(np.float32)
    f = (x - (r - magic).astype(np.float32)).astype(np.float32)
    p = (np.float32(0.054592825) * f + np.float32(0.24221784)).astype(np.float32)
    p = (p * f + np.float32(0.6933686)).astype(np.float32)
    p = (p * f + np.float32(1.0)).astype(np.float32)
    bits = (p.view(np.int32).astype(np.int64) + (r.view(np.int32).astype(np.int64) << 23)) & 0xFFFFFFFF
    return bits.astype(np.uint32).view(np.float32)


def test_ex2_poly_accuracy_and_edges():
    x = np.linspace(-100.0, 8.0, 400001).astype(np.float32)
    y, ref = ex2_poly(x), np.exp2(x.astype(np.float64))
    assert np.all(np.isfinite(y)) and np.all(y > 0)
    assert float(np.max(np.abs(y / ref - 1))) < 1.3e-4  # << 2^-9, the bf16 rounding of the probabilities
    ints = np.arange(-100, 9, dtype=np.float32)
    assert np.array_equal(ex2_poly(ints), np.exp2(ints).astype(np.float32))  # f = 0 -> exactly 2^i
    # monotone (the online softmax only needs p <= 2^8 and order preservation up to rounding)
    assert np.all(np.diff(ex2_poly(np.linspace(-20, 8, 100001).astype(np.float32))) >= -1e-9)
    tiny = ex2_poly(np.array([-126.0, -200.0, -1e30], dtype=np.float32))
    assert np.all(tiny < 1e-37)  # clamped: contributes nothing to a row sum
