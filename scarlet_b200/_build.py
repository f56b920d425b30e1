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


def source_hash():
    """sha256 over the CUDA sources and headers: compiled into the library (``sb_source_hash``) and written next to it, so
    that a stale binary is rebuilt instead of being loaded against a changed ABI (mtimes do not survive a snapshot copy)."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES) + sorted(HEADERS):
        h.update(open(os.path.join(SRC_DIR, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "scarlet_b200.h"), "rb").read())
    return h.hexdigest()[:16]


def up_to_date():
    try:
        return os.path.exists(LIB) and open(LIB + ".hash").read().strip() == source_hash()
    except OSError:
        return False


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
    digest = source_hash()
    if not force and not up_to_date() and os.path.exists(LIB):
        force = True  # the sources changed under an existing binary (or it predates the hash file)
    flags = FLAGS + ['-DSB_SRC_HASH="%s"' % digest]
    jobs = []
    for f in SOURCES:
        src, obj = os.path.join(SRC_DIR, f), os.path.join(OBJ_DIR, f[:-3] + ".o")
        if force or _newer(obj, [src] + _deps()):
            jobs.append([NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
    if jobs:
        with ThreadPoolExecutor(len(jobs)) as pool:
            list(pool.map(subprocess.check_call, jobs))
    objs = [os.path.join(OBJ_DIR, f[:-3] + ".o") for f in SOURCES]
    if jobs or _newer(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcufft"])
    with open(LIB + ".hash", "w") as f:
        f.write(digest + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
