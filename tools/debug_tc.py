"""GPU diagnostic for the tcgen05 GEMM modes (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mr_gan_b200.engine import FoldGroup

fg = FoldGroup([(16, 100, 20, 1)], precision="tf32")
rng = np.random.default_rng(0)


def ref(mode, A, B):
    A, B = A.astype(np.float64), B.astype(np.float64)
    return A @ B if mode == 0 else (A @ B.T if mode == 1 else A.T @ B)


def shapes(mode, M, N, K):
    if mode == 0: return (M, K), (K, N)
    if mode == 1: return (M, K), (N, K)
    return (K, M), (K, N)


for mode in (1, 0, 2):
    for (M, N, K) in ((50, 128, 32), (50, 128, 64), (50, 200, 101), (150, 1000, 1201), (300, 40, 40), (1201, 1000, 150) if mode == 2 else (150, 250, 251)):
        sa, sb = shapes(mode, M, N, K)
        A, B = rng.standard_normal(sa).astype(np.float32), rng.standard_normal(sb).astype(np.float32)
        C = fg.debug_gemm(mode, A, B, use_tc=True)
        Cs = fg.debug_gemm(mode, A, B, use_tc=False)
        R = ref(mode, A, B)
        scale = np.abs(R).max()
        print("mode %d M=%d N=%d K=%d: tc err %.3e  simt err %.3e" % (mode, M, N, K, np.abs(C - R).max() / scale, np.abs(Cs - R).max() / scale))
    # structure probe: small integers are exact in tf32
    M, N, K = 50, 64, 16
    sa, sb = shapes(mode, M, N, K)
    if mode == 2:
        M, N, K = 40, 64, 16
        sa, sb = shapes(mode, M, N, K)
    # A picks contraction index r % K for output row r; B holds unique codes
    A = np.zeros(sa, np.float32); B = np.zeros(sb, np.float32)
    if mode == 0:
        for r in range(M): A[r, r % K] = 1
        for k in range(K):
            for n in range(N): B[k, n] = k * 64 + n          # code: k*256 + n
    elif mode == 1:
        for r in range(M): A[r, r % K] = 1
        for n in range(N):
            for k in range(K): B[n, k] = k * 64 + n
    else:
        for m in range(M): A[m % K, m] = 1
        for k in range(K):
            for n in range(N): B[k, n] = k * 64 + n
    C = fg.debug_gemm(mode, A, B, use_tc=True)
    R = ref(mode, A, B)
    bad = np.argwhere(np.abs(C - R) > 0.5)
    print("  probe: %d / %d wrong" % (len(bad), C.size))
    for (r, n) in bad[:12]:
        c = C[r, n]
        print("   out[%d,%d] = %.1f (k=%d n=%d)  want k=%d n=%d" % (r, n, c, int(c) // 64, int(c) % 64, int(R[r, n]) // 64, int(R[r, n]) % 64))
fg.close()
