"""Stage-1 check of the fp16-operand (kind::f16) variant of k_gemm_tc: the three operand-major combinations of the
step (forward: A MN-major / B K-major, dX: both K-major, dW: both MN-major) against numpy on fp16-rounded inputs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mr_gan_b200.engine import FoldGroup

rng = np.random.default_rng(0)
ok = True
with FoldGroup([(36, 100, 50, 1)], precision="tf32") as fg:
    for name, mode, shp in [("fwd", 0, (150, 1201, 1000)), ("dx", 1, (150, 1000, 1201)), ("dw", 2, (150, 1201, 1000)),
                            ("fwd small", 0, (50, 101, 500)), ("dw small", 2, (100, 251, 250))]:
        if mode == 0:
            M, K, N = shp; A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N)) * 0.05
        elif mode == 1:
            M, K, N = shp; A = rng.standard_normal((M, K)) * 0.01; B = rng.standard_normal((N, K)) * 0.05
        else:
            K, M, N = shp; A = rng.standard_normal((K, M)); B = rng.standard_normal((K, N)) * 0.01
        A = A.astype(np.float32); B = B.astype(np.float32)
        Ah = A.astype(np.float16).astype(np.float64); Bh = B.astype(np.float16).astype(np.float64)
        want = Ah @ Bh if mode == 0 else (Ah @ Bh.T if mode == 1 else Ah.T @ Bh)
        for tc, tag in ((2, "f16"), (1, "tf32")):
            try:
                got = fg.lib.mrgan_debug_gemm  # noqa: F841  (symbol present)
                out = np.zeros(want.shape, dtype=np.float32)
                from mr_gan_b200 import _lib
                fg._chk(fg.lib.mrgan_debug_gemm(fg._h, mode, want.shape[0], want.shape[1], shp[1] if mode != 2 else shp[0],
                                                _lib.fptr(np.ascontiguousarray(A)), _lib.fptr(np.ascontiguousarray(B)), _lib.fptr(out), tc))
                err = np.linalg.norm(out - want) / np.linalg.norm(want)
                print("%-10s %-5s rel Frobenius error %.3e  max|diff| %.3e" % (name, tag, err, np.abs(out - want).max()))
                if tc == 2 and not err < 2e-5:
                    ok = False
            except Exception as e:  # a trap / launch failure surfaces here
                print("%-10s %-5s FAILED: %s" % (name, tag, e))
                ok = False
                if tc == 2:
                    raise
print("F16 GEMM", "OK" if ok else "MISMATCH")
