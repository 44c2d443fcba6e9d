"""ORACLE (test infrastructure only) -- restatement of mr_gan.py's epoch loop
(mr_gan.py:183-234) and fold preparation (mr_gan.py:86-107) on top of the step oracle.

PARITY UNPINNED (see gan_oracle.py header).  The loop takes the epoch index
arrays as INPUTS (the reference builds them with unseeded numpy permutations,
mr_gan.py:189-202) and replays the device noise stream (oracle/philox.py), so
that the CUDA epoch path and this loop see identical batches and noise.
Never imported by the product path.
"""
import numpy as np

from . import gan_oracle as O
from . import philox


def d_noise(key, step, rows, D, row0, tids=philox.TID_D_LAYER):
    """The 5 GaussianNoise draws of one D application on stacked rows [row0, row0+rows)."""
    widths = (D,) + O.D_WIDTHS[:4]
    return [philox.normal(key, step, tids[l], rows, widths[l], row0=row0) for l in range(5)]


def d_transforms(key, step, rows, D, row0, dropout_rate, tids=philox.TID_D_LAYER):
    """Dropout variant (others/wganlpctsemi.py:170) of ``d_noise``: N(0,1) for the input layer, Dropout keep factors for
    the four transforms in front of hidden layers 2..5 -- the ``noise`` list of gan_oracle.disc_forward(dropout=True)."""
    widths = (D,) + O.D_WIDTHS[:4]
    return [philox.normal(key, step, tids[0], rows, widths[0], row0=row0)] + \
           [philox.dropout_factor(key, step, tids[l], rows, widths[l], dropout_rate, row0=row0) for l in range(1, 5)]


def prep_fold(X_train, X_test, y_train, y_test, percentlabeled, percentunlabeled, rng, K=O.K_CLASSES):
    """mr_gan.py:96-107: StandardScaler, shuffle, first 10*percent rows per class as labeled.

    Returns scaled train/test plus ROW INDICES (into the shuffled train set) of
    the labeled subset and (table 6) of the unlabeled subset.
    """
    mu = X_train.mean(axis=0)
    sd = X_train.std(axis=0)
    sd = np.where(sd == 0.0, 1.0, sd)               # sklearn's zero-variance guard
    Xtr, Xte = (X_train - mu) / sd, (X_test - mu) / sd
    perm = rng.permutation(len(Xtr))                # sklearn.utils.shuffle
    Xtr, ytr = Xtr[perm], y_train[perm]
    nlab = int(10 * percentlabeled)
    lab_rows = np.concatenate([np.nonzero(ytr == j)[0][:nlab] for j in range(K)])
    unl_rows = None
    if percentunlabeled is not None:
        nun = nlab + int(10 * percentunlabeled)
        unl_rows = np.concatenate([np.nonzero(ytr == j)[0][:nun] for j in range(K)])
    return Xtr, Xte, ytr, y_test, lab_rows, unl_rows


def tiled_perm(rng, n_total, n_sub):
    """mr_gan.py:189 (and :197-201): floor(N/L) permutations of L plus one of N mod L."""
    parts = [rng.permutation(n_sub) for _ in range(n_total // n_sub)]
    parts.append(rng.permutation(n_total % n_sub))
    return np.concatenate(parts).astype(np.int64)


def epoch_indices(rng, n_train, lab_rows, unl_rows=None):
    """Row indices (into X_train) of the 3 streams used by one epoch, mr_gan.py:189-202.

    The reference also draws a third unlabeled permutation that it never uses
    (trainx_unl3, :195/:202); it is drawn here too so the generator state advances alike.
    """
    idx_lab = lab_rows[tiled_perm(rng, n_train, len(lab_rows))]
    if unl_rows is None:
        u1, u2, _ = (rng.permutation(n_train) for _ in range(3))
    else:
        u1, u2, _ = (unl_rows[tiled_perm(rng, n_train, len(unl_rows))] for _ in range(3))
    return idx_lab.astype(np.int32), np.asarray(u1, np.int32), np.asarray(u2, np.int32)


def device_perm(key, epoch, stream, tile, n):
    """A permutation of range(n) as the device draws it (csrc/kernels_simt.cuh:k_epoch_perm): the order of the 64-bit keys
    (philox(i, tile, epoch, 0x50 + stream).x << 32) | i, ascending."""
    x = philox.philox4x32_10(np.arange(n, dtype=np.int64), np.uint64(tile), np.uint64(epoch & 0xFFFFFFFF), np.uint64(0x50 + stream),
                             key[0], key[1])[0]
    keys = (np.broadcast_to(x, (n,)).astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    return np.argsort(keys, kind="stable")


def device_epoch_indices(key, epoch, n_train, lab_rows, unl_rows=None):
    """mr_gan.py:189-202 with the device's permutations: the three index streams of mrgan_train_epoch_seeded."""
    def tiled(stream, rows, L):
        parts = [device_perm(key, epoch, stream, j, L) for j in range(n_train // L)]
        if n_train % L:
            parts.append(device_perm(key, epoch, stream, n_train // L, n_train % L))
        p = np.concatenate(parts)
        return (rows[p] if rows is not None else p).astype(np.int32)
    lab_rows = np.asarray(lab_rows)
    out = [tiled(0, lab_rows, len(lab_rows))]
    for s in (1, 2):
        out.append(tiled(s, None if unl_rows is None else np.asarray(unl_rows), n_train if unl_rows is None else len(unl_rows)))
    return out


def train_epoch(model, X_train, y_train, idx_lab, idx_unl, idx_unl2, key, rng_step, B=O.BATCH_GAN):
    """mr_gan.py:204-217 with the device noise stream.  Returns per-step stats [nb,4] and the new rng_step.

    ``rng_step`` counts executed train steps (D and G steps alike); it is the
    Philox ``step`` word.  Stacked-row convention of the device: D step rows =
    [labeled | unlabeled | fake]; G step rows = [fake | real].
    """
    D = X_train.shape[1]
    nb = X_train.shape[0] // B
    stats = np.zeros((nb, 4))
    for t in range(nb):
        sl = slice(t * B, (t + 1) * B)
        z = philox.normal(key, rng_step, philox.TID_Z, B, O.NOISE_SIZE)
        ll, lu, te = model.disc_step(
            X_train[idx_lab[sl]], y_train[idx_lab[sl]], X_train[idx_unl[sl]], z,
            d_noise(key, rng_step, B, D, 0), d_noise(key, rng_step, B, D, B), d_noise(key, rng_step, B, D, 2 * B))
        rng_step += 1
        z = philox.normal(key, rng_step, philox.TID_Z, B, O.NOISE_SIZE)
        lg = model.gen_step(X_train[idx_unl2[sl]], z,
                            d_noise(key, rng_step, B, D, 0), d_noise(key, rng_step, B, D, B))
        rng_step += 1
        stats[t] = (ll, lu, te, lg)
    return stats, rng_step


def eval_batches(model, X_test, y_test, B=O.BATCH_GAN):
    """mr_gan.py:219-223: mean of per-batch errors over the first floor(N/B) batches."""
    nb = X_test.shape[0] // B
    return float(np.mean([model.test_batch(X_test[t * B:(t + 1) * B], y_test[t * B:(t + 1) * B]) for t in range(nb)]))
