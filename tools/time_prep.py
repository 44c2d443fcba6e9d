import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from mr_gan_b200.engine import FoldGroup
from mr_gan_b200.model import fold_key
G = 74
X, y, folds = bench.make_jobs(2, G, seed=0)
ntr, nte = len(folds[0][0].train_rows), len(folds[0][0].test_rows)
fg = FoldGroup([(1200, ntr, nte, fold_key(0, i)) for i in range(G)], precision="tf32")
for rep in range(3):
    t0 = time.perf_counter(); fg.load_dataset(0, X, y); t1 = time.perf_counter()
    for i, (f, rng) in enumerate(folds):
        fg.prepare_fold(i, 0, f.train_rows, f.test_rows)
    t2 = time.perf_counter()
    print("rep %d: load_dataset %.1f ms, 74 x prepare_fold %.1f ms" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    if rep == 0:
        idx = np.stack([np.arange(ntr, dtype=np.int32)] * G)
        t0 = time.perf_counter(); fg.train_epoch(idx, idx, idx); print("first epoch (graph build) %.1f ms" % (1e3 * (time.perf_counter() - t0)))
        t0 = time.perf_counter(); fg.train_epoch(idx, idx, idx); print("second epoch %.1f ms" % (1e3 * (time.perf_counter() - t0)))
fg.close()
