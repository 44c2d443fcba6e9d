"""A/B of the fused epilogue heads (MRGAN_HEADS bit mask): device ms of one D step / G step over 74 folds."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    from mr_gan_b200.engine import FoldGroup
    G, D = 74, 1200
    rng = np.random.default_rng(0)
    with FoldGroup([(D, 100, 50, i + 1) for i in range(G)], precision=sys.argv[1]) as fg:
        X = rng.standard_normal((100, D)).astype(np.float32); y = (np.arange(100) % 6).astype(np.int32)
        for i in range(G):
            fg.load_fold(i, X, y, X[:50], y[:50])
        print("heads=%s %s disc %.4f gen %.4f" % (os.environ.get("MRGAN_HEADS"), sys.argv[1], fg.time_op("disc_step", reps=5), fg.time_op("gen_step", reps=5)))
else:
    for mask in (0, 1, 2, 4, 8, 10, 15):
        subprocess.run([sys.executable, __file__, "f16"], env=dict(os.environ, MRGAN_HEADS=str(mask)))
