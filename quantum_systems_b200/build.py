"""Build libqsb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: ``python -m quantum_systems_b200.build [--force]``.  The driver's "does it build" check calls
this through ``__graft_entry__.build()``; nvcc cross-compiles without a GPU.
"""

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libqsb200.so")

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
]
# The CUDA runtime is linked dynamically (libcudart.so.12, the one torch has already loaded): the shipped library
# then carries none of the runtime's own entry points or strings.
LINK_FLAGS = ["-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
OBJ_DIR = os.path.join(PKG_DIR, "build")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG_DIR, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > built for d in deps)


def _compile(nvcc, src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    deps = [src] + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG_DIR, "..", "include", "*.h"))
    if os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
        return obj, ""
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return obj, proc.stderr


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` (one nvcc process per file, in parallel) and link ``libqsb200.so``.
    Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libqsb200.so cannot be built and there is no CPU fallback")
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for old in glob.glob(os.path.join(OBJ_DIR, "*.o")):
            os.remove(old)
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(lambda src: _compile(nvcc, src, verbose), sources()))
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc] + LINK_FLAGS + ["-o", tmp] + [obj for obj, _ in results]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
