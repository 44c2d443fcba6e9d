"""ORACLE (test infrastructure only) -- counter-based Gaussian noise replay.

This file is part of the CPU oracle.  It is imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg ONLY; the product path (``mr_gan_b200``) never imports it.

What it restates
----------------
The reference draws its noise on the host with numpy's global Mersenne
Twister (``np.random.normal`` mr_gan.py:206,212) and inside Theano's
``GaussianNoise`` layers (mr_gan.py:118-126), and it is deliberately
unseeded (mr_gan.py:74-75), so there is NO reference noise stream to match.
The B200 path therefore defines its own replayable stream -- Philox4x32-10
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
published Random123 algorithm) followed by Box-Muller -- and this file is the
numpy restatement of exactly that definition:

    ctr  = (row >> 2, col, step, tensor_id)        key = (key0, key1)
    x[4] = philox4x32_10(ctr, key)
    pair = (row & 3) >> 1 ; xa, xb = x[2*pair], x[2*pair+1]
    u1 = ((xa >> 9) + 0.5) * 2**-23 ; u2 = ((xb >> 9) + 0.5) * 2**-23
    rad = sqrt(-2 ln u1) ; ang = pi * (2*u2 - 1)
    n(row, col) = rad * cos(ang)  if row is even else  rad * sin(ang)

Pinned by the Random123 known-answer vectors (tests/test_oracle_philox.py).
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

# tensor ids of the noise streams (shared with mr_gan_b200/csrc/common.cuh)
TID_D_LAYER = (0, 1, 2, 3, 4)   # GaussianNoise in front of D's 5 hidden Dense layers
TID_Z = 5                       # generator input z


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def fold_key(seed, fold):
    """Per-fold Philox key from a 64-bit user seed (host side does the same)."""
    s = (int(seed) + 0x9E3779B97F4A7C15 * (int(fold) + 1)) & 0xFFFFFFFFFFFFFFFF
    # splitmix64 finaliser
    s ^= s >> 30
    s = (s * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    s ^= s >> 27
    s = (s * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    s ^= s >> 31
    return s & 0xFFFFFFFF, (s >> 32) & 0xFFFFFFFF


def normal(key, step, tensor_id, rows, cols, row0=0, dtype=np.float64):
    """N(0,1) matrix [rows, cols] of the device stream (see module docstring).

    ``row0`` is the stacked-row index of the first row (the device indexes noise
    by the row's position inside the stacked D batch).
    """
    r = (np.arange(rows, dtype=np.int64) + row0)[:, None]
    c = np.arange(cols, dtype=np.int64)[None, :]
    x = philox4x32_10(r >> 2, c, np.uint64(step & 0xFFFFFFFF), np.uint64(tensor_id), key[0], key[1])
    x = [np.broadcast_to(v, (rows, cols)) for v in x]
    pair = ((r & 3) >> 1).astype(bool)
    pair = np.broadcast_to(pair, (rows, cols))
    xa = np.where(pair, x[2], x[0]).astype(np.uint64)
    xb = np.where(pair, x[3], x[1]).astype(np.uint64)
    u1 = ((xa >> np.uint64(9)).astype(np.float64) + 0.5) * 2.0 ** -23
    u2 = ((xb >> np.uint64(9)).astype(np.float64) + 0.5) * 2.0 ** -23
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = np.pi * (2.0 * u2 - 1.0)
    odd = np.broadcast_to((r & 1).astype(bool), (rows, cols))
    return np.where(odd, rad * np.sin(ang), rad * np.cos(ang)).astype(dtype)


def dropout_factor(key, step, tensor_id, rows, cols, rate, row0=0, dtype=np.float64):
    """Dropout(rate) keep factors [rows, cols] of the device stream (csrc/common.cuh:dropout4): the same counters as
    ``normal``; element (row, col) uses word ``row & 3`` of philox(row >> 2, col, step, tensor_id):
    u = ((x >> 9) + 0.5) * 2**-23 (computed in float32 like the device), kept iff u >= rate, factor 1 / (1 - rate)."""
    r = (np.arange(rows, dtype=np.int64) + row0)[:, None]
    c = np.arange(cols, dtype=np.int64)[None, :]
    x = philox4x32_10(r >> 2, c, np.uint64(step & 0xFFFFFFFF), np.uint64(tensor_id), key[0], key[1])
    x = np.stack([np.broadcast_to(v, (rows, cols)) for v in x])
    w = np.take_along_axis(x, np.broadcast_to((r & 3)[None], (1, rows, cols)), axis=0)[0]
    u = ((w >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(rate))
    return np.where(u >= np.float32(rate), inv, np.float32(0.0)).astype(dtype)
