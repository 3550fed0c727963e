"""Numerical probe of the int8-sliced (Ozaki-style) fp64 GEMM used by the tcgen05 MLP kernel: emulates the exact integer
arithmetic of the kernel in numpy and compares with an mpmath-free high-precision reference (float128 via longdouble)."""
import sys
import numpy as np

rng = np.random.default_rng(0)
K, N, B = 128, 548, 2000
W = rng.normal(0, 0.1, size=(N, K)) + rng.uniform(-1, 1, size=(N, K)) / np.sqrt(128) / 1000
bias = rng.normal(0, 1, size=N)
h = np.tanh(rng.normal(0, 1.5, size=(B, K)))

ref = (h.astype(np.longdouble) @ W.T.astype(np.longdouble)).astype(np.float64)
f64 = h @ W.T


def slices(v_int, n_slices, bits=8):
    """two's complement decomposition of int64 v (|v| < 2^(bits*n_slices-1)): top slice signed, the rest unsigned"""
    out = []
    half = 1 << (bits - 1)
    for s in range(n_slices):                    # balanced digits in [-2^(bits-1), 2^(bits-1)), least significant first
        d = ((v_int + half) & ((1 << bits) - 1)) - half
        out.append(d)
        v_int = (v_int - d) >> bits
    assert np.all(v_int == 0)
    return out[::-1]                             # most significant first


def ozaki(h, W, n_slices, n_levels, bits=8):
    T = bits * n_slices
    # A: fixed point with T-1 fractional bits (|h| < 1)
    hi = np.rint(h * 2.0 ** (T - 2)).astype(np.int64)
    hi = np.clip(hi, -(2 ** (T - 2)), 2 ** (T - 2))
    # B: per output column power-of-two scale so that |W/scale| < 1
    e = np.ceil(np.log2(np.abs(W).max(axis=1) * (1 + 2.0 ** -40)))
    sc = 2.0 ** e
    wi = np.rint(W / sc[:, None] * 2.0 ** (T - 2)).astype(np.int64)
    wi = np.clip(wi, -(2 ** (T - 2)), 2 ** (T - 2))
    hs, ws = slices(hi, n_slices, bits), slices(wi, n_slices, bits)
    acc = [np.zeros((h.shape[0], W.shape[0]), dtype=np.int64) for _ in range(n_levels)]
    n_mma = 0
    for p in range(n_slices):
        for q in range(n_slices):
            if p + q < n_levels:
                acc[p + q] += hs[p] @ ws[q].T
                n_mma += 1
    assert max(np.abs(a).max() for a in acc) < 2 ** 31
    # value = sum_l acc[l] * 2^(-bits*l) * 2^(2*bits*(n_slices-1)) / 2^(2(T-1))
    v = np.zeros_like(acc[0], dtype=np.float64)
    for l in range(n_levels - 1, -1, -1):
        v = v * 2.0 ** -bits + acc[l].astype(np.float64)
    v = v * 2.0 ** (2 * bits * (n_slices - 1) - 2 * (T - 2))
    return v * sc[None, :], n_mma


scale = np.abs(ref).max()
print("fp64 GEMM vs longdouble: max abs err %.2e (rel to max |out| %.2e)" % (np.abs(f64 - ref).max(), np.abs(f64 - ref).max() / scale))
for n_slices, n_levels in [(5, 5), (6, 5), (6, 6), (6, 7), (7, 6), (7, 7), (7, 8)]:
    v, n = ozaki(h, W, n_slices, n_levels)
    err = np.abs(v - ref).max()
    print("slices %d levels %d: %2d MMAs  max abs err %.2e  (rel %.2e)" % (n_slices, n_levels, n, err, err / scale))
