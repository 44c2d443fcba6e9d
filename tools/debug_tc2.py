import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mr_gan_b200.engine import FoldGroup
fg = FoldGroup([(16, 100, 20, 1)], precision="tf32")
rng = np.random.default_rng(0)
for (M, N, K) in ((150, 250, 6), (150, 250, 8), (150, 250, 4), (150, 250, 12), (150, 250, 250), (150, 500, 250), (150, 1000, 500), (50, 100, 1000)):
    A, B = rng.standard_normal((M, K)).astype(np.float32), rng.standard_normal((N, K)).astype(np.float32)
    C = fg.debug_gemm(1, A, B, use_tc=True)
    R = A.astype(np.float64) @ B.astype(np.float64).T
    e = np.abs(C - R) / np.abs(R).max()
    print("dx M=%d N=%d K=%d err %.3e; worst rows %s cols %s" % (M, N, K, e.max(), np.unique(np.argwhere(e > 0.01)[:, 0])[:10], np.unique(np.argwhere(e > 0.01)[:, 1])[:10]))
fg.close()
