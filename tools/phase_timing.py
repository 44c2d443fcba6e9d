"""Per-phase timestamps of the forward GEMM of D layer 1 (debug build with -DMRGAN_PHASE_TIMING, MRGAN_LIB=...)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mr_gan_b200.engine import FoldGroup
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
G, D = 74, 1200
rng = np.random.default_rng(0)
with FoldGroup([(D, 100, 50, i + 1) for i in range(G)], precision=prec) as fg:
    X = rng.standard_normal((100, D)).astype(np.float32); y = (np.arange(100) % 6).astype(np.int32)
    for i in range(G):
        fg.load_fold(i, X, y, X[:50], y[:50])
    print("fwd1 ms", fg.time_op("fwd1", reps=2))
    print("dw1 ms", fg.time_op("dw1", reps=2))
