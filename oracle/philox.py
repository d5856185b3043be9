"""Philox4x32-10 in numpy (TEST INFRASTRUCTURE): the published counter-based generator (Salmon et al., SC'11; multipliers 0xD2511F53 /
0xCD9E8D57, Weyl keys 0x9E3779B9 / 0xBB67AE85) as dcs_dropout uses it — counter = (offset + i / 4, 0, 0, 0), key = seed, element
i takes word i % 4, keep iff word >= p * 2^32 — so the kernel's mask stream has a CPU statement to be checked against."""
import numpy as np, torch
def philox4x32_10(counter_lo, counter_hi, key_lo, key_hi):
    c0 = counter_lo.astype(np.uint64); c1 = counter_hi.astype(np.uint64); c2 = np.zeros_like(c0); c3 = np.zeros_like(c0)
    k0, k1 = np.uint64(key_lo), np.uint64(key_hi)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for r in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask; k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c0, c1, c2, c3
def dropout_mask(n, p, seed, offset):
    n4 = (n + 3) // 4
    c = np.arange(n4, dtype=np.uint64) + np.uint64(offset)
    r = philox4x32_10(c & np.uint64(0xFFFFFFFF), c >> np.uint64(32), seed & 0xFFFFFFFF, seed >> 32)
    bits = np.stack(r, 1).reshape(-1)[:n]
    thr = np.uint64(min(np.float32(p) * np.float32(4294967296.0), np.float32(4294967295.0)))
    keep = bits >= thr
    return torch.from_numpy(keep.astype(np.float32)) * np.float32(1.0 / (1.0 - p))
