"""Philox4x32-10 on the host (numpy), bit-identical to the kernels' generator (csrc/ag_device.cuh philox4x32_10 /
philox_uniform2): lets a caller reproduce the actions (stream 0) and reset candidates (stream 1) any environment drew
inside a rollout launch, so in-kernel actions need not be shipped back.  Pure index arithmetic -- not a compute path."""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """counter (c0..c3) and key (k0, k1): uint32 arrays of one shape -> four uint32 arrays"""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32).copy() for x in (c0, c1, c2, c3))
    k0, k1 = np.asarray(k0, dtype=np.uint32).copy(), np.asarray(k1, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + _W0
            k1 = k1 + _W1
    return c0, c1, c2, c3


def uniform2(seed, gid, draw, stream):
    """two float64 uniforms in [0,1) with 53 random bits each (numpy's rand() recipe: (a>>5, b>>6)), for the counter
    (gid, draw, stream) and key `seed`"""
    gid = np.asarray(gid, dtype=np.uint64)
    draw = np.asarray(draw, dtype=np.uint64)
    seed = np.uint64(seed)
    shape = np.broadcast(gid, draw).shape
    gid, draw = np.broadcast_to(gid, shape), np.broadcast_to(draw, shape)
    w = philox4x32_10((gid & _MASK).astype(np.uint32), (gid >> np.uint64(32)).astype(np.uint32),
                      (draw & _MASK).astype(np.uint32), np.full(shape, stream, dtype=np.uint32),
                      np.full(shape, int(seed) & 0xFFFFFFFF, dtype=np.uint32), np.full(shape, int(seed) >> 32, dtype=np.uint32))
    u0 = ((w[0] >> np.uint32(5)).astype(np.float64) * 67108864.0 + (w[1] >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0
    u1 = ((w[2] >> np.uint32(5)).astype(np.float64) * 67108864.0 + (w[3] >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0
    return u0, u1
