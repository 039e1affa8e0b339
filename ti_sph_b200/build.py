"""Builds libtisph.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The library is kept at ti_sph_b200/lib/libtisph.so together with a hash of the sources it was
built from, so that a prebuilt copy travels to GPU boxes and is rebuilt only when stale.
"""
import hashlib
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "lib")
LIB = os.path.join(LIBDIR, "libtisph.so")
SOURCES = ["tisph.cu"]
HEADERS = ["tisph_kernels.cuh", "tisph_device.cuh", "tisph_walk.cuh", "tisph_lists.cuh", "tisph_shard.cuh", "tisph_gen1.cuh", "tisph_voxel.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
if os.environ.get("TISPH_CHECKS") == "1":       # debug build with device-side bounds checks (tisph_device.cuh)
    NVCC_FLAGS = NVCC_FLAGS + ["-DTISPH_CHECKS"]


def _source_hash():
    h = hashlib.sha256()
    for f in [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "tisph.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    try:
        with open(LIB + ".hash") as fh:
            return os.path.exists(LIB) and fh.read().strip() == _source_hash()
    except OSError:
        return False


def have_nvcc():
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


def build(force=False, verbose=False):
    """Compile the library if it is missing or stale. Returns its path."""
    if not force and is_fresh():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libtisph.so (there is no CPU fallback)")
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(LIB + ".hash", "w") as fh:
        fh.write(_source_hash())
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
