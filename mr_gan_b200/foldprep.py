"""Host-side fold preparation and epoch permutations of the reference, seedable.

mr_gan.py:86-107 (split, StandardScaler, shuffle, labeled / unlabeled subsets) and
mr_gan.py:189-202 (tiled permutations).  The reference uses numpy's unseeded global RNG
(mr_gan.py:74-75); here every draw comes from an explicit ``numpy.random.Generator`` so runs
are reproducible.  The device receives ROW INDICES into the resident X_train instead of the
gathered copies the reference builds every epoch."""
from collections import namedtuple

import numpy as np
from sklearn import preprocessing
from sklearn.model_selection import train_test_split

FoldData = namedtuple("FoldData", "x_train y_train x_test y_test lab_rows unl_rows")


def prepare_fold(X, y, percentlabeled, percentunlabeled=None, trainTestSets=None, rng=None, n_classes=6):
    rng = rng if rng is not None else np.random.default_rng()
    test_ratio = 200 * n_classes                                  # mr_gan.py:81
    num_labeled = int(10 * percentlabeled)                        # mr_gan.py:82
    if trainTestSets is None:                                     # mr_gan.py:87-88
        X_train, X_test, y_train, y_test = train_test_split(
            X, y, test_size=test_ratio, stratify=y, random_state=int(rng.integers(2 ** 31)))
    else:
        X_train, X_test, y_train, y_test = trainTestSets          # mr_gan.py:90
    scaler = preprocessing.StandardScaler()                       # mr_gan.py:96-98
    X_train = scaler.fit_transform(np.asarray(X_train, dtype=np.float64))
    X_test = scaler.transform(np.asarray(X_test, dtype=np.float64))
    y_train, y_test = np.asarray(y_train), np.asarray(y_test)
    perm = rng.permutation(len(X_train))                          # sklearn.utils.shuffle, mr_gan.py:101
    X_train, y_train = X_train[perm], y_train[perm]
    lab_rows = np.concatenate([np.nonzero(y_train == j)[0][:num_labeled] for j in range(n_classes)])   # :102
    unl_rows = None
    if percentunlabeled is not None:                              # mr_gan.py:106-107
        n_unl = num_labeled + int(10 * percentunlabeled)
        unl_rows = np.concatenate([np.nonzero(y_train == j)[0][:n_unl] for j in range(n_classes)])
    return FoldData(X_train.astype(np.float32), y_train.astype(np.int32), X_test.astype(np.float32),
                    y_test.astype(np.int32), lab_rows, unl_rows)


def tiled_perm(rng, n_total, n_sub):
    """mr_gan.py:189: floor(N/L) permutations of L rows followed by a permutation of N mod L."""
    parts = [rng.permutation(n_sub) for _ in range(n_total // n_sub)]
    parts.append(rng.permutation(n_total % n_sub))
    return np.concatenate(parts).astype(np.int64)


def epoch_indices(rng, n_train, lab_rows, unl_rows=None):
    """Row indices into X_train of trainx / trainx_unl / trainx_unl2 (mr_gan.py:189-202).

    The third unlabeled permutation the reference draws and never uses (trainx_unl3) is drawn
    and discarded so the generator advances exactly as the reference's does."""
    idx_lab = lab_rows[tiled_perm(rng, n_train, len(lab_rows))]
    if unl_rows is None:
        u1, u2, _ = (rng.permutation(n_train) for _ in range(3))
    else:
        u1, u2, _ = (unl_rows[tiled_perm(rng, n_train, len(unl_rows))] for _ in range(3))
    return idx_lab.astype(np.int32), np.asarray(u1, np.int32), np.asarray(u2, np.int32)


FoldIndex = namedtuple("FoldIndex", "train_rows test_rows y_train y_test lab_rows unl_rows")


def prepare_fold_indices(y, train_idx, test_idx, percentlabeled, percentunlabeled=None, rng=None, n_classes=6):
    """Index-only version of ``prepare_fold`` for the device-side fold preparation (``FoldGroup.prepare_fold``):
    the same shuffle and labeled / unlabeled subset selection (mr_gan.py:101-107), but X never leaves the GPU.
    Draws from ``rng`` exactly like ``prepare_fold(trainTestSets=...)`` does, so both paths pick identical rows."""
    rng = rng if rng is not None else np.random.default_rng()
    y = np.asarray(y)
    train_idx, test_idx = np.asarray(train_idx), np.asarray(test_idx)
    perm = rng.permutation(len(train_idx))
    train_rows = train_idx[perm]
    y_train = y[train_rows]
    num_labeled = int(10 * percentlabeled)
    lab_rows = np.concatenate([np.nonzero(y_train == j)[0][:num_labeled] for j in range(n_classes)])
    unl_rows = None
    if percentunlabeled is not None:
        n_unl = num_labeled + int(10 * percentunlabeled)
        unl_rows = np.concatenate([np.nonzero(y_train == j)[0][:n_unl] for j in range(n_classes)])
    return FoldIndex(train_rows.astype(np.int32), test_idx.astype(np.int32), y_train.astype(np.int32),
                     y[test_idx].astype(np.int32), lab_rows, unl_rows)
