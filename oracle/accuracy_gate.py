"""ORACLE (test infrastructure only) -- the north star's accuracy gate: final fold accuracy of the CUDA path within
+-0.5 pt of the ORACLE's over a fixed seed set.

A GAN fold-training is a chaotic trajectory: two runs of the SAME implementation that differ only in their noise draws
end 1.8 pt apart (std, measured with this very set-up), so a 0.5 pt gate on a handful of folds is a coin toss.  The gate
is therefore statistical: N_FOLDS = 120 independent fold-trainings (40 data seeds x 3 splits) of a reduced problem
(synthetic MREO-shape temperature channel at 1 s: D = 100; 1200 training rows, 40 % labeled; 2400 test rows; B = 50,
10 epochs; class signal raised so that the accuracy sits near 80 %, neither chance nor saturated), compared through their
MEAN accuracy: std of the mean difference = 1.8 / sqrt(120) = 0.16 pt, i.e. 0.5 pt is a 3-sigma band.  Data, splits,
scaler, labeled subset, initial weights and epoch permutations are identical on both sides (built here); the noise
streams are each side's own (the device's Philox stream vs torch.randn), which is exactly what "over a seed set" allows.

The oracle side (torch-CPU fp32 twin of gan_oracle.py, ~8 min on 4 threads) is run ONCE in the development container
and committed as tests/golden/accuracy_gate.npz:      python -m oracle.accuracy_gate
PARITY UNPINNED (see gan_oracle.py header).  Never imported by the product path.
"""
import os
import sys
import time

import numpy as np

from . import fold_loop, gan_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "accuracy_gate.npz")
N_SEEDS, N_SPLITS, EPOCHS, BATCH = 40, 3, 10, 50
PERCENT_LABELED = 8            # 80 labeled rows per class of the 200 available (mr_gan.py:82)
CLASS_AMPLITUDE = 0.45
POKES = 50                     # 6 materials x 12 objects x 50 pokes = 3600 rows: 1200 train / 2400 test per split


def dataset(seed):
    from mr_gan_b200 import synthetic
    keep = synthetic.CLASS_AMPLITUDE
    synthetic.CLASS_AMPLITUDE = CLASS_AMPLITUDE
    try:
        return synthetic.synthetic_dataset(1, forcetempTime=1, pokes=POKES, seed=seed)
    finally:
        synthetic.CLASS_AMPLITUDE = keep


def fold_cases(seeds=range(N_SEEDS)):
    """Yields every fold of the gate: scaled data, labeled rows, initial weights, per-epoch row indices."""
    from sklearn.model_selection import StratifiedKFold
    for seed in seeds:
        X, y = dataset(seed)
        skf = StratifiedKFold(N_SPLITS, shuffle=True, random_state=seed)
        for k, (big, small) in enumerate(skf.split(X, y)):          # train on the small part, test on the rest
            rng = np.random.default_rng([seed, k])
            Xtr, Xte, ytr, yte, lab, _ = fold_loop.prep_fold(X[small], X[big], y[small], y[big], PERCENT_LABELED, None, rng)
            D = X.shape[1]
            pD = [p.astype(np.float32) for p in O.init_disc_params(D, rng)]
            pG = [p.astype(np.float32) for p in O.init_gen_params(D, rng)]
            idx = [fold_loop.epoch_indices(rng, len(Xtr), lab) for _ in range(EPOCHS)]
            yield dict(seed=seed, k=k, Xtr=Xtr.astype(np.float32), Xte=Xte.astype(np.float32), ytr=ytr.astype(np.int32),
                       yte=yte.astype(np.int32), pD=pD, pG=pG, idx=idx)


def oracle_accuracy(c, noise_seed=0):
    import torch
    from . import torch_twin as T
    torch.manual_seed(1000003 * noise_seed + 97 * c['seed'] + c['k'])
    m = T.TorchGan(c['pD'], c['pG'])
    Xtr, ytr = c['Xtr'], c['ytr']
    for il, iu, iu2 in c['idx']:
        for t in range(len(Xtr) // BATCH):
            sl = slice(t * BATCH, (t + 1) * BATCH)
            m.disc_step(Xtr[il[sl]], ytr[il[sl]], Xtr[iu[sl]], torch.randn(BATCH, O.NOISE_SIZE))
            m.gen_step(Xtr[iu2[sl]], torch.randn(BATCH, O.NOISE_SIZE))
    return 1.0 - m.test_batch(c['Xte'], c['yte'])


def main():
    import torch
    torch.set_num_threads(int(os.environ.get("ORACLE_THREADS", "4")))
    acc, t0 = [], time.time()
    for c in fold_cases():
        acc.append(oracle_accuracy(c))
        print("seed %d split %d accuracy %.4f  (%.0f s)" % (c['seed'], c['k'], acc[-1], time.time() - t0), flush=True)
    acc = np.array(acc)
    np.savez_compressed(OUT, acc=acc, n_seeds=N_SEEDS, n_splits=N_SPLITS, epochs=EPOCHS, percent=PERCENT_LABELED,
                        class_amplitude=CLASS_AMPLITUDE, pokes=POKES)
    print("mean accuracy %.4f over %d folds, fold std %.4f -> %s" % (acc.mean(), len(acc), acc.std(), OUT))


if __name__ == "__main__":
    sys.exit(main())
