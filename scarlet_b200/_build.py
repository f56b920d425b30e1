"""Build the CUDA library in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC_DIR = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libscarlet_b200.so")
SOURCES = ["scarlet_b200.cu", "spectral_f32.cu", "spectral_f64.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "fft_core.cuh", "spectral.cuh"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-I" + os.path.join(ROOT, "include")]


def _deps():
    return [os.path.join(SRC_DIR, f) for f in HEADERS] + [os.path.join(ROOT, "include", "scarlet_b200.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile ``csrc/*.cu`` -> ``lib/libscarlet_b200.so`` (cross-compiles without a GPU); translation units in parallel."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = []
    for f in SOURCES:
        src, obj = os.path.join(SRC_DIR, f), os.path.join(OBJ_DIR, f[:-3] + ".o")
        if force or _newer(obj, [src] + _deps()):
            jobs.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
    if jobs:
        with ThreadPoolExecutor(len(jobs)) as pool:
            list(pool.map(subprocess.check_call, jobs))
    objs = [os.path.join(OBJ_DIR, f[:-3] + ".o") for f in SOURCES]
    if jobs or _newer(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcufft"])
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
