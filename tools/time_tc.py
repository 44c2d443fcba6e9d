import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mr_gan_b200.engine import FoldGroup
fg = FoldGroup([(16, 100, 20, 1)], precision="tf32")
print("mode M N K groups -> us")
for mode in (1, 0, 2):
    for (M, N, K, G) in ((150, 128, 32, 1), (150, 128, 320, 1), (150, 128, 3200, 1), (150, 128, 3200, 12), (150, 1000, 1201, 12), (150, 250, 251, 12), (50, 128, 3200, 1)):
        if mode == 2:
            M, N, K = K, N, M      # dW shapes: C[M in, N out], contraction = rows
        us = fg.debug_gemm_time(mode, M, N, K, G, reps=20) * 1e3
        print(mode, M, N, K, G, "-> %.1f us" % us)
fg.close()
