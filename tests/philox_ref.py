"""numpy restatements of the counter-based generators of SWK_MODE_FAST (spinwalk_b200/csrc/walk_fast.cuh): Philox4x32-10 and
Philox2x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) and the kernel's Box-Muller bit
layout.  Test infrastructure: tests/test_rng.py pins them on the known-answer vectors of the Random123 distribution (kat_vectors),
tests/test_rng_gpu.py compares the device code with them word for word."""
import numpy as np

M4_0, M4_1, W_0, W_1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
M2 = 0xD256D193
FIXED_KEY = (0x243F6A88, 0x85A308D3)  # philox_fixed: the key of the displacement stream (the seed lives in the counter)
_U32 = np.uint64(0xFFFFFFFF)


def philox4x32(ctr, key, rounds=10):
    """ctr [n,4] uint32, key (k0, k1) -> [n,4] uint32"""
    c = [np.asarray(ctr)[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(rounds):
        p0, p1 = np.uint64(M4_0) * c[0], np.uint64(M4_1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & _U32, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & _U32]
        k0, k1 = (k0 + W_0) & 0xFFFFFFFF, (k1 + W_1) & 0xFFFFFFFF
    return np.stack(c, 1).astype(np.uint32)


def philox2x32(ctr, key, rounds=10):
    """ctr [n,2] uint32, key [n] uint32 (or scalar) -> [n,2] uint32"""
    c0, c1 = (np.asarray(ctr)[:, i].astype(np.uint64) for i in range(2))
    k = np.broadcast_to(np.asarray(key, np.uint64), c0.shape).copy()
    for _ in range(rounds):
        p = np.uint64(M2) * c0
        c0, c1 = (p >> np.uint64(32)) ^ k ^ c1, p & _U32
        k = (k + np.uint64(W_0)) & _U32
    return np.stack([c0, c1], 1).astype(np.uint32)


def normals6(block):
    """walk_fast.cuh normals6_fast in float64: block [n,4] uint32 -> [n,6] (even attempt x y z, odd attempt x y z).
    pair i: radius uniform = top 23 bits of word i (u in (0,1]), angle = 19 bits: word i [8:0] ++ a 10-bit field of word 3."""
    b = np.asarray(block).astype(np.uint64)
    out = np.empty((b.shape[0], 6))
    for i in range(3):
        w = b[:, i]
        field = (b[:, 3] >> np.uint64(22 - 10 * i)) & np.uint64(0x3FF)
        u = 1.0 - (w >> np.uint64(9)).astype(np.float64) / 2.0**23
        m19 = ((w & np.uint64(0x1FF)) << np.uint64(10)) | field
        t = 2.0 * np.pi * (1.0 + m19.astype(np.float64) / 2.0**19)
        r = np.sqrt(-2.0 * np.log(u))
        out[:, 2 * i], out[:, 2 * i + 1] = r * np.cos(t), r * np.sin(t)
    return out
