"""numpy restatement of the Philox4x32-10 stream and of the sampler's transforms - TEST INFRASTRUCTURE ONLY.

Philox4x32-10 is the counter-based generator of Salmon et al. (Random123, SC'11); the round function and the
constants below are the published ones, and `KAT` holds Random123's known-answer vectors for it. The CUDA sampler
(csrc/philox.cu) must reproduce the raw stream bit-exactly; the float transforms (Box-Muller, Poisson inversion)
are compared within float tolerance / statistically.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# (counter[4], key[2]) -> output[4], from Random123's kat_vectors for philox4x32-10
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(x, np.uint64) for x in (c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def raw_stream(n_groups, seed, offset):
    """uint32 [n_groups, 4]: group g = Philox(counter = offset + g, key = seed) as csrc/philox.cu lays it out."""
    ctr = np.arange(n_groups, dtype=np.uint64) + np.uint64(offset)
    out = philox4x32_10(ctr & MASK, ctr >> np.uint64(32), np.zeros_like(ctr), np.zeros_like(ctr),
                        np.uint64(seed) & MASK, np.uint64(seed) >> np.uint64(32))
    return np.stack(out, 1).astype(np.uint32)


def u01(x):
    return ((x >> 8).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def normals(n_groups, seed, offset):
    r = raw_stream(n_groups, seed, offset)
    out = np.empty((n_groups, 4), np.float32)
    for a, b, i in ((0, 1, 0), (2, 3, 2)):
        rad = np.sqrt(np.float32(-2.0) * np.log(u01(r[:, a])))
        th = np.float32(6.28318530717958647692) * u01(r[:, b])
        out[:, i] = rad * np.cos(th)
        out[:, i + 1] = rad * np.sin(th)
    return out


FACTOR_KEY = 0x5bd1e9955bd1e995


def normal_demand(B, S, T, mean, std, rho, clip, seed, offset):
    """[T,S,B] float32, the value csrc/philox.cu assigns to element (t,s,b)."""
    n = T * S * B
    z = normals((n + 3) // 4, seed, offset).reshape(-1)[:n].reshape(T, S, B)
    if rho > 0:
        nf = T * B
        f = normals((nf + 3) // 4 + 1, seed ^ FACTOR_KEY, offset).reshape(-1)
        fac = f[np.arange(nf)].reshape(T, 1, B)
        z = np.float32(np.sqrt(np.float32(rho))) * fac + np.float32(np.sqrt(np.float32(1 - rho))) * z
    d = np.asarray(mean, np.float32)[None, :, None] + np.asarray(std, np.float32)[None, :, None] * z
    return np.maximum(d, 0) if clip else d
