"""Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
SC'11; Random123 reference implementation), vectorised in numpy.  TEST INFRASTRUCTURE: the checker of the in-kernel
disturbance generator of `dtmpc_plant_step_*` (include/dtmpc.h, struct dtmpc_disturbance).  Pinned by the
known-answer vectors of Random123's kat_vectors file (tests/test_oracle_golden.py).

The reference draws its disturbance from torch's global generator (core/systems/dubins.py:57-65):
w = low + (high - low) * U[0,1)^3.  The product keeps the distribution and replaces the stream by a counter-based one,
keyed by (seed, global instance, closed-loop step), so that a draw does not depend on batch size or sharding.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr [...,4] uint32, key [...,2] uint32 (broadcastable) -> [...,4] uint32."""
    c = [np.asarray(ctr)[..., j].astype(np.uint64) for j in range(4)]
    k0 = np.asarray(key)[..., 0].astype(np.uint64)
    k1 = np.asarray(key)[..., 1].astype(np.uint64)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & MASK]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def disturbance(seed, first_instance, n, step, low, high, dtype=np.float64):
    """w [n,3] exactly as plant_step_kernel draws it (dtmpc.cu): counter = (instance lo, instance hi, step, draw)."""
    gi = np.uint64(first_instance) + np.arange(n, dtype=np.uint64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    ctr = np.zeros((n, 4), dtype=np.uint32)
    ctr[:, 0], ctr[:, 1], ctr[:, 2] = (gi & MASK).astype(np.uint32), (gi >> np.uint64(32)).astype(np.uint32), np.uint32(step)
    a = philox4x32_10(ctr, key)
    ctr[:, 3] = 1
    b = philox4x32_10(ctr, key)
    low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
    if dtype == np.float64:
        u64 = lambda lo, hi: (((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        u = np.stack([u64(a[:, 0], a[:, 1]), u64(a[:, 2], a[:, 3]), u64(b[:, 0], b[:, 1])], axis=1)
    else:
        u32 = lambda lo: ((lo >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))
        u = np.stack([u32(a[:, 0]), u32(a[:, 2]), u32(b[:, 0])], axis=1)
    return (low + (high - low) * u).astype(dtype)
