"""The arithmetic the single kernels rely on, emulated in numpy (no GPU): an fp32 value split into fp16
hi + lo keeps 22 significant bits while the lo part is a normal fp16 (|x| >= 2^-3; below that the error is
absolute, 2^-25: which is why the WEIGHTS are pre-scaled by a power of two and the activations need not be), so
products of split operands accumulated in fp32 match an fp32 FMA chain -
the three-product form of the tcgen05 layers (ah*bh + ah*bl + al*bh) and the four-product form of the
mma.sync layers ((W_hi + W_lo) . (h_hi + h_lo)), with the per-layer power-of-two weight scale of
remora_b200/csrc/rb200_mega.cu (pow2_scale: largest |w| just under 2^15)."""
import numpy as np


def split16(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def pow2_scale(w):
    mx = float(np.abs(w).max())
    _, e = np.frexp(mx)
    return float(np.ldexp(1.0, 15 - int(e)))


def test_split_keeps_22_bits():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * rng.choice([1e-3, 1.0, 50.0], size=100000)).astype(np.float32)
    hi, lo = split16(x)
    rel = np.abs((hi + lo).astype(np.float64) - x) / np.maximum(np.abs(x), 1e-30)
    big = np.abs(x) >= 0.125                      # lo part >= 2^-14: a normal fp16
    assert rel[big].max() < 2.0 ** -21
    assert np.abs((hi + lo) - x)[~big].max() <= 2.0 ** -24   # lo in fp16's subnormal range: absolute error


def _fp32_dot(a_terms, b_terms):
    """sum over K of the products of the given operand pairs, every product and every addition rounded to fp32
    (the products of fp16 values are exact in fp32; the tensor core accumulates in fp32)."""
    acc = np.zeros(a_terms[0].shape[:-1] + b_terms[0].shape[1:], dtype=np.float32)
    for a, b in zip(a_terms, b_terms):
        for k in range(a.shape[-1]):
            acc = (acc + np.float32(1) * (a[..., k:k + 1] * b[k:k + 1, ...]).astype(np.float32)).astype(np.float32)
    return acc


def test_three_and_four_product_forms_match_fp32():
    rng = np.random.default_rng(1)
    K, M, N = 640, 16, 8                          # merge_conv1: 5 taps x 128 channels
    a = (rng.standard_normal((M, K)) * 0.7).astype(np.float32)          # activations (after swish: O(1))
    w = (rng.standard_normal((K, N)) * 0.08).astype(np.float32)         # weights
    exact = a.astype(np.float64) @ w.astype(np.float64)
    scale = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64)  # sum of |terms|
    s = pow2_scale(w)
    ah, al = split16(a)
    wh, wl = split16((w * np.float32(s)).astype(np.float32))
    three = _fp32_dot([ah, ah, al], [wh, wl, wh]) / np.float32(s)
    four = _fp32_dot([ah, ah, al, al], [wh, wl, wh, wl]) / np.float32(s)
    fp32 = _fp32_dot([a], [w])
    err = lambda y: float((np.abs(y.astype(np.float64) - exact) / scale).max())  # noqa: E731
    assert err(fp32) < 5e-7
    assert err(three) < 6e-7 and err(four) < 6e-7           # as good as the fp32 chain ...
    single = _fp32_dot([ah], [wh]) / np.float32(s)
    assert err(single) > 20 * err(three)                     # ... which one fp16 product per MAC is not


def test_recurrence_operands_stay_in_fp16_range():
    """h is in (-1, 1): its hi part is a normal or subnormal fp16 with absolute error < 2^-25 after the lo part;
    W_hh scaled to just under 2^15 leaves the accumulated gate sums (64 terms) far below fp32 overflow."""
    rng = np.random.default_rng(2)
    h = np.tanh(rng.standard_normal(4096)).astype(np.float32) * rng.choice([1.0, 1e-3, 1e-6], size=4096).astype(np.float32)
    hi, lo = split16(h)
    assert np.abs((hi + lo) - h).max() <= 2.0 ** -24
    w = (rng.standard_normal((256, 64)) * 0.6).astype(np.float32)
    s = pow2_scale(w)
    assert 2.0 ** 14 <= np.abs(w * s).max() < 2.0 ** 15
    assert np.isfinite((w * np.float32(s)).astype(np.float16)).all()
