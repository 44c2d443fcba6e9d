"""Small driver for ncu: a fold group at the bench shapes (D=1200, B=50) but few batches per epoch."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mr_gan_b200.engine import FoldGroup
from mr_gan_b200.model import init_disc, init_gen, fold_key

ap = argparse.ArgumentParser()
ap.add_argument("--folds", type=int, default=48)
ap.add_argument("--precision", default="tf32")
ap.add_argument("--D", type=int, default=1200)
ap.add_argument("--n-train", type=int, default=300)
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--batch", type=int, default=50)
a = ap.parse_args()
rng = np.random.default_rng(0)
G, D, ntr, nte = a.folds, a.D, a.n_train, 100
fg = FoldGroup([(D, ntr, nte, fold_key(0, i)) for i in range(G)], precision=a.precision, batch=a.batch, eval_each_epoch=a.batch <= 256)
X = rng.standard_normal((ntr, D)).astype(np.float32); y = (np.arange(ntr) % 6).astype(np.int32)
pD, pG = init_disc(D, rng), init_gen(D, rng)
for i in range(G):
    fg.set_params(i, 0, pD); fg.set_params(i, 1, pG)
    fg.load_fold(i, X, y, X[:nte], y[:nte])
idx = np.stack([rng.permutation(ntr) for _ in range(G)]).astype(np.int32)
for e in range(a.epochs):
    st = fg.train_epoch(idx, idx, idx)
    print("epoch", e, "ms", fg.last_device_ms, "launches", fg.kernel_launches, st[0])
fg.close()
