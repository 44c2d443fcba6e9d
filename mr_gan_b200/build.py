"""Builds libmrgan.so (the C-ABI of include/mrgan.h) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmrgan.so")
SOURCES = ["mrgan_api.cu"]
HEADERS = ["common.cuh", "kernels_simt.cuh", "kernels_tc.cuh", os.path.join("..", "..", "include", "mrgan.h")]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libmrgan.so cannot be built (there is no CPU fallback)")
    return p


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> mr_gan_b200/libmrgan.so.  Returns the library path."""
    if not force and not _stale():
        return LIB
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
             "-Xcompiler", "-fPIC", "-shared", "-ldl", "--use_fast_math=false"]
    flags = [f for f in flags if f != "--use_fast_math=false"]
    if os.path.exists(os.path.join(CSRC, "kernels_tc.cuh")):
        flags += ["-DMRGAN_WITH_TC", "-lcuda"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    cmd = [nvcc_path()] + flags + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed (exit %d): %s" % (r.returncode, " ".join(cmd)))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
