"""ORACLE -- test infrastructure only.  numpy restatement of the counter-based normal generator of the sampling path
(`jf_normal_rows`, jammy_flows_b200/csrc/rng.cuh): Philox4x32-10 (Salmon et al., SC'11; the published Random123
known-answer vectors are checked in tests/test_philox_oracle.py) keyed by the seed, counter = (global row, pair index),
two 53-bit uniforms per block, Box-Muller.  It replaces the reference's host numpy RNG
(main/default.py:1661-1668: `numpy.random.normal` + H2D copy); only `tests/` may import it."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32"""
    c = [np.asarray(ctr[..., i], dtype=np.uint32) for i in range(4)]
    k0, k1 = np.asarray(key[..., 0], dtype=np.uint32), np.asarray(key[..., 1], dtype=np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def normal_rows(seed, first_row, n_rows, dim):
    """[n_rows, dim] float64 standard normals; row i depends only on (seed, first_row + i)."""
    rows = np.arange(first_row, first_row + n_rows, dtype=np.uint64)
    n_pairs = (dim + 1) // 2
    out = np.empty((n_rows, 2 * n_pairs), dtype=np.float64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    for j in range(n_pairs):
        ctr = np.stack([(rows & MASK).astype(np.uint32), (rows >> np.uint64(32)).astype(np.uint32),
                        np.full(n_rows, j, dtype=np.uint32), np.zeros(n_rows, dtype=np.uint32)], axis=-1)
        r = philox4x32_10(ctr, np.broadcast_to(key, (n_rows, 2))).astype(np.uint64)
        # 53-bit uniforms in (0, 1): ((hi << 21) ^ (lo >> 11)) + 0.5) * 2^-53
        u1 = ((((r[:, 0] << np.uint64(21)) ^ (r[:, 1] >> np.uint64(11))) & np.uint64((1 << 53) - 1)).astype(np.float64) + 0.5) * 2.0 ** -53
        u2 = ((((r[:, 2] << np.uint64(21)) ^ (r[:, 3] >> np.uint64(11))) & np.uint64((1 << 53) - 1)).astype(np.float64) + 0.5) * 2.0 ** -53
        rad = np.sqrt(-2.0 * np.log(u1))
        out[:, 2 * j] = rad * np.cos(2.0 * np.pi * u2)
        out[:, 2 * j + 1] = rad * np.sin(2.0 * np.pi * u2)
    return out[:, :dim]
