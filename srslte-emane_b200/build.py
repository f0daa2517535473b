"""In-tree build of the CUDA library (libsrslte_b200.so) for sm_100a with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU-only dev container; the resulting .so is
git-ignored but travels to the GPU box with the repo snapshot.  Every source is compiled to its own
object (in parallel, only when it or a header is newer) and the objects are linked into the library.
"""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsrslte_b200.so")

SOURCES = ["tdec_kernels.cu", "frontend_kernels.cu", "capi.cu", "lte_tables.cpp"]
HEADERS = ["tdec_kernels.h", "lte_tables.h", os.path.join("..", "..", "include", "srslte_b200.h"),
           os.path.join("..", "..", "include", "srslte_b200_compat.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _headers_mtime():
    return max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)


def _obj(src):
    return os.path.join(OBJ, os.path.splitext(src)[0] + ".o")


def _stale():
    t = _mtime(LIB)
    return t == 0.0 or any(_mtime(os.path.join(CSRC, s)) > t for s in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile csrc/ into libsrslte_b200.so if it is missing or older than its sources."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_mtime()
    todo = [s for s in SOURCES
            if force or _mtime(_obj(s)) < max(_mtime(os.path.join(CSRC, s)), hdr)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", _obj(src)]
        r = subprocess.run(cmd, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=max(1, len(todo))) as ex:
        results = list(ex.map(compile_one, todo))
    for src, rc, out in results:
        if out and (verbose or rc):
            print(out)
        if rc:
            raise subprocess.CalledProcessError(rc, f"nvcc -c {src}")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [_obj(s) for s in SOURCES] + ["-o", LIB],
                          cwd=HERE)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
