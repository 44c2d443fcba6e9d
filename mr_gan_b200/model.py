"""Model construction of mr_gan.py:109-133 / mr_nn.py:101-113, host side: shapes, Keras-style
initialisers (Glorot-uniform kernels, zero biases, BN gamma=1/beta=0) and per-fold RNG keys."""
import numpy as np

NOISE_SIZE = 100                       # mr_gan.py:77
D_WIDTHS = (1000, 500, 250, 250, 250)  # mr_gan.py:119-127
G_HIDDEN = 500                         # mr_gan.py:111,113
MATERIALS = ['plastic', 'glass', 'fabric', 'metal', 'wood', 'ceramic']   # mr_gan.py:80


def disc_shapes(D, K=len(MATERIALS)):
    dims = (D,) + D_WIDTHS + (K,)
    out = []
    for i in range(6):
        out += [(dims[i], dims[i + 1]), (dims[i + 1],)]
    return out


def gen_shapes(D, noise_dim=NOISE_SIZE):
    return [(noise_dim, G_HIDDEN), (G_HIDDEN,), (G_HIDDEN,), (G_HIDDEN,),
            (G_HIDDEN, G_HIDDEN), (G_HIDDEN,), (G_HIDDEN, D), (D,)]


def _glorot_uniform(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_disc(D, rng, K=len(MATERIALS)):
    """Dense layers of mr_gan.py:117-128 with Keras default initialisers."""
    return [_glorot_uniform(rng, s) if len(s) == 2 else np.zeros(s, np.float32) for s in disc_shapes(D, K)]


def init_gen(D, rng, noise_dim=NOISE_SIZE):
    """Dense/BN layers of mr_gan.py:110-114 with Keras default initialisers."""
    p = [_glorot_uniform(rng, s) if len(s) == 2 else np.zeros(s, np.float32) for s in gen_shapes(D, noise_dim)]
    p[2] = np.ones(G_HIDDEN, np.float32)
    return p


def fold_key(seed, fold):
    """64-bit Philox key of one fold's noise streams (splitmix64 of seed and fold number)."""
    m = 0xFFFFFFFFFFFFFFFF
    s = (int(seed) + 0x9E3779B97F4A7C15 * (int(fold) + 1)) & m
    s ^= s >> 30
    s = (s * 0xBF58476D1CE4E5B9) & m
    s ^= s >> 27
    s = (s * 0x94D049BB133111EB) & m
    s ^= s >> 31
    return s
