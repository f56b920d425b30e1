"""CPU: the in-register FFT building blocks of the fused spectral kernels (scarlet_b200/csrc/fft_core.cuh) against
numpy.fft, through a g++-built harness that emulates the lanes of the two-stage transform (tests/fftcore/)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from numpy.testing import assert_allclose

from conftest import ROOT

LENGTHS = [(6, 8), (8, 8), (8, 9), (8, 10), (8, 12), (10, 10), (10, 12), (8, 16), (12, 12), (10, 15), (10, 16), (12, 15), (12, 16),
           (10, 20), (12, 18), (15, 16), (16, 16), (16, 18), (15, 20), (16, 20), (18, 18), (18, 20), (16, 24), (20, 20)]


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "fftcore", "fftcore_host.cpp")
    core = os.path.join(ROOT, "scarlet_b200", "csrc", "fft_core.cuh")
    out = os.path.join(ROOT, "tests", "fftcore", "_build", "libfftcore_host.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(core)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-I" + os.path.dirname(core), "-o", out, src])
    lib = ctypes.CDLL(out)
    for name in ("fftcore_two_stage_f32", "fftcore_two_stage_f64", "fftcore_reg_f32", "fftcore_reg_f64"):
        getattr(lib, name).restype = ctypes.c_int
    return lib


def _run(fn, args, x, dtype):
    a = np.ascontiguousarray(np.stack([x.real, x.imag], axis=-1), dtype=dtype)
    out = np.zeros_like(a)
    assert fn(*args, a.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)) == 0
    return out[..., 0].astype(np.float64) + 1j * out[..., 1].astype(np.float64)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 18, 20, 24, 25])
def test_register_fft_vs_numpy(harness, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for inv, ref in ((0, np.fft.fft(x)), (1, np.fft.ifft(x) * n)):
        f64 = _run(lambda a, o: harness.fftcore_reg_f64(n, a, o, inv), (), x, np.float64)
        assert_allclose(f64, ref, atol=1e-14 * n)
        f32 = _run(lambda a, o: harness.fftcore_reg_f32(n, a, o, inv), (), x, np.float32)
        assert_allclose(f32, ref, atol=3e-6)


@pytest.mark.parametrize("r1,r2", LENGTHS)
def test_two_stage_fft_vs_numpy(harness, r1, r2):
    n = r1 * r2
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for inv, ref in ((0, np.fft.fft(x)), (1, np.fft.ifft(x) * n)):
        f64 = _run(lambda a, o: harness.fftcore_two_stage_f64(r1, r2, a, o, inv), (), x, np.float64)
        assert_allclose(f64, ref, atol=2e-13 * np.sqrt(n))
        f32 = _run(lambda a, o: harness.fftcore_two_stage_f32(r1, r2, a, o, inv), (), x, np.float32)
        err = np.abs(f32 - ref).max() / np.abs(ref).max()
        assert err < 6e-7, err   # float32 accuracy on a par with cuFFT (tools/fft_accuracy_probe.py: ~3e-7 max)
