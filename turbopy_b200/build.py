"""Build libturbopy_b200.so (hand-written sm_100a kernels + the C ABI) in-tree
with nvcc.  No torch C++ extension machinery: the library has a plain C ABI
(include/turbopy_b200.h) and is loaded with ctypes.

    python -m turbopy_b200.build [--force] [--verbose]
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OBJ_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(HERE, "libturbopy_b200.so")

SOURCES = ["capi.cu", "push_kernels.cu", "fd_kernels.cu", "reduce_kernels.cu", "scan_kernels.cu", "field_kernels.cu",
           "sort_kernels.cu", "host_path.cu", "nccl_link.cu"]

# -fmad=false is part of the arithmetic contract (no FMA contraction: results
# are bit-identical to the reference's numpy); -lineinfo maps ncu's source page
# back to these files.
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", INCLUDE,
]
# tuning experiments only (e.g. TURBOPY_B200_NVCC_FLAGS="-DTP_MIN_BLOCKS=3")
NVCC_FLAGS += os.environ.get("TURBOPY_B200_NVCC_FLAGS", "").split()


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libturbopy_b200.so")
    return exe


def _digest():
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, INCLUDE):
        for name in sorted(os.listdir(root)):
            with open(os.path.join(root, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    return h.hexdigest()


def needs_build():
    stamp = os.path.join(OBJ_DIR, "stamp")
    if not (os.path.exists(LIB_PATH) and os.path.exists(stamp)):
        return True
    with open(stamp) as f:
        return f.read().strip() != _digest()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns the
    library path; the ptxas resource report goes to _build/ptxas.log."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stdout + r.stderr
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{logs[src]}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB_PATH, *objs, "-ldl",
                        "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "w") as f:
        for src in SOURCES:
            f.write(f"==== {src}\n{logs[src]}\n")
    with open(os.path.join(OBJ_DIR, "stamp"), "w") as f:
        f.write(_digest())
    if verbose:
        for src in SOURCES:
            print(f"==== {src}\n{logs[src]}")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
