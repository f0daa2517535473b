"""In-tree build of the CUDA library (libsrslte_b200.so) for sm_100a with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU-only dev container; the resulting .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsrslte_b200.so")

SOURCES = ["tdec_kernels.cu", "frontend_kernels.cu", "capi.cu", "lte_tables.cpp"]
HEADERS = ["tdec_kernels.h", "lte_tables.h", os.path.join("..", "..", "include", "srslte_b200.h"),
           os.path.join("..", "..", "include", "srslte_b200_compat.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, s))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/ into libsrslte_b200.so if it is missing or older than its sources."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", LIB]
    subprocess.check_call(cmd, cwd=HERE)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
