import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mr_gan_b200.engine import FoldGroup
fg = FoldGroup([(16, 100, 20, 1)], precision="tf32")
print("mode M(rows) N(feat) K groups -> us")
for mode in (0, 1):
    for (M, N, K, G) in ((150, 1000, 1201, 48), (100, 1000, 1201, 48), (50, 1000, 1201, 48), (16, 1000, 1201, 48), (150, 500, 1001, 48), (50, 500, 1001, 48), (150, 250, 501, 48), (150, 1000, 1201, 24), (150, 1000, 1201, 96)):
        us = fg.debug_gemm_time(mode, M, N, K, G, reps=20) * 1e3
        kb = (K + 31) // 32
        bn = (M + 15) // 16 * 16
        ctas = ((N + 127) // 128) * G
        tot = ctas * kb * (16384 + bn * 128)
        print(mode, M, N, K, G, "-> %.1f us   ctas %d  kb %d  SM-ingest %.0f MB -> %.2f TB/s; weights %.0f MB -> %.2f TB/s" % (us, ctas, kb, tot / 1e6, tot / us / 1e6, N * K * 4 * G / 1e6, N * K * 4 * G / us / 1e6))
fg.close()
