#!/usr/bin/env python
"""float32 device convolution error vs float64 NumPy at several FFT grid sizes (diagnostic; GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scarlet_b200 import fft as sfft  # noqa: E402

rng = np.random.default_rng(0)
for N, P, grids in ((128, 21, (138, 140, 144, 150, 160, 192, 256)), (256, 41, (276, 280, 288, 300, 320, 384, 512))):
    img = (rng.random((5, N, N)) * np.exp(-((np.mgrid[:N, :N] - N / 2) ** 2).sum(0) / (N / 4) ** 2)).astype(np.float32)
    ker = rng.random((5, P, P))
    ker /= ker.sum(axis=(1, 2))[:, None, None]
    for F in grids:
        pad = np.zeros((5, F, F))
        for i in range(P):
            for j in range(P):
                pad[:, (i - P // 2) % F, (j - P // 2) % F] = ker[:, i, j]
        khat = np.fft.rfftn(pad, axes=(1, 2))
        big = np.zeros((5, F, F))
        big[:, :N, :N] = img
        ref = np.fft.irfftn(np.fft.rfftn(big, axes=(1, 2)) * khat, (F, F), axes=(1, 2))[:, :N, :N]
        out = sfft.device_convolve(img, khat, (F, F))
        err = np.abs(out - ref)
        print("N=%d P=%d F=%d: max err / max %.3e   rms err / rms %.3e" % (N, P, F, err.max() / np.abs(ref).max(),
              np.sqrt((err ** 2).mean()) / np.sqrt((ref ** 2).mean())), flush=True)
