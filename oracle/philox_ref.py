"""TEST INFRASTRUCTURE ONLY -- host restatement of the counter-based noise stream the kernels generate under
GPKL_FLAG_PHILOX_EPS (gp-vae_b200/csrc/gpkl_common.cuh, "counter-based N(0,1) noise"): Philox4x32-10 exactly as published
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123; the generator behind cuRAND's
curandStatePhilox4_32_10_t) followed by cuRAND's Box-Muller of curand_normal4.  It stands in for tf.random_normal inside
tf_kernel (/root/reference/src/Models/Full_GP_VAE_dynamic_time.py:166), whose stream is not reproducible anyway.
Pinned by the Random123 known-answer vectors (tests/test_philox_cpu.py)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    for _ in range(10):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        with np.errstate(over="ignore"):
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, -1)


def philox_normal(seed, n):
    """The first n draws of the stream: element e is lane e & 3 of Philox(counter = (lo32(e>>2), hi32(e>>2), 0, 0),
    key = (lo32(seed), hi32(seed))), normals pairwise by Box-Muller in float32 like curand_normal4."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    g = np.arange((n + 3) // 4, dtype=np.uint64)
    ctr = np.stack([(g & MASK).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                    np.zeros_like(g, dtype=np.uint32), np.zeros_like(g, dtype=np.uint32)], -1)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32), (len(g), 2))
    r = philox4x32_10(ctr, key)
    f = np.float32
    c = f(2.3283064e-10)
    c2pi = f(c * f(6.2831855))
    out = np.empty((len(g), 4), dtype=np.float32)
    for a in (0, 2):
        u = (r[:, a].astype(np.float32) * c + c / f(2.0)).astype(np.float32)
        v = (r[:, a + 1].astype(np.float32) * c2pi + c2pi / f(2.0)).astype(np.float32)
        s = np.sqrt(f(-2.0) * np.log(u)).astype(np.float32)
        out[:, a] = s * np.sin(v).astype(np.float32)
        out[:, a + 1] = s * np.cos(v).astype(np.float32)
    return out.reshape(-1)[:n]
