"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) restated in numpy
from the published algorithm, pinned by the Random123 known-answer vectors (tests/test_cpu.py). Test infrastructure."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = (int(x) & MASK for x in ctr)
    k0, k1 = (int(x) & MASK for x in key)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox_uniforms(seed, pixel, sample, n):
    """u(seed, pixel, sample, dimension i) as the library defines it: key = (seed lo, seed hi), counter =
    (pixel, sample, i // 4, 0), uniform = (word >> 8) * 2^-24."""
    out = np.zeros(n, np.float32)
    for i in range(n):
        w = philox4x32_10((pixel, sample, i // 4, 0), (seed & MASK, (seed >> 32) & MASK))[i % 4]
        out[i] = np.float32(w >> 8) * np.float32(1.0 / 16777216.0)
    return out
