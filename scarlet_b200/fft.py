"""Padding / centring conventions, fast FFT shapes and PSF matching (host setup) + device convolution.

Mirrors the slice of scarlet/fft.py the fitting path relies on: ``_centered`` 9-36, ``_pad`` 82-113,
``_get_fft_shape`` 116-167, ``Fourier`` 170-313 (image + cached transforms), ``match_psf`` 334-365,
``convolve`` 368-396.  ``match_psf`` runs once per observation on the host (as in the reference,
renderer.py:198-202); ``convolve`` -- the per-iteration operation -- runs on the GPU (cuFFT) through
``sb_fft_convolve_*`` with the kernel transform K^ computed once here.
"""
import os

import numpy as np
from scipy import fftpack

from . import _native as nat


def _centered(arr, newshape):
    """Centre crop of ``arr`` to ``newshape`` (start index ``(old-new+1)//2``)."""
    newshape = np.asarray(newshape)
    cur = np.array(arr.shape)
    if not np.all(newshape <= cur):
        raise ValueError("arr must be larger than newshape in every dimension")
    lo = (cur - newshape + 1) // 2
    return arr[tuple(slice(int(a), int(a + n)) for a, n in zip(lo, newshape))]


def _pad(arr, newshape, axes=None):
    """Zero-pad to ``newshape`` along ``axes``; odd excess puts the extra line in front (centre-right rule)."""
    arr = np.asarray(arr)
    if axes is None:
        axes = tuple(range(arr.ndim))
    elif np.isscalar(axes):
        axes = (int(axes),)
    widths = [(0, 0)] * arr.ndim
    for n, ax in enumerate(axes):
        extra = int(newshape[n]) - arr.shape[ax]
        if extra < 0:
            raise ValueError("newshape must not be smaller than the array")
        widths[ax] = ((extra + 1) // 2, extra // 2)
    return np.pad(arr, widths, mode="constant")


def _get_fft_shape(im_or_shape1, im_or_shape2, padding=3, axes=None, max=False):
    """5-smooth FFT lengths for a linear convolution; last axis even, second-to-last even if operand 2 is."""
    s1 = np.asarray(im_or_shape1.shape if hasattr(im_or_shape1, "shape") else im_or_shape1)
    s2 = np.asarray(im_or_shape2.shape if hasattr(im_or_shape2, "shape") else im_or_shape2)
    if len(s1) != len(s2):
        raise ValueError("img1 and img2 must have the same number of dimensions")
    if axes is None:
        axes = range(len(s1))
    elif np.isscalar(axes):
        axes = [axes]
    tot = [np.maximum(s1[a], s2[a]) if max else s1[a] + s2[a] for a in axes]
    fast = [int(fftpack.next_fast_len(int(t + padding))) for t in tot]
    while fast[-1] % 2:
        fast[-1] = int(fftpack.next_fast_len(fast[-1] + 1))
    if len(fast) > 1 and s2[-2] % 2 == 0:
        while fast[-2] % 2:
            fast[-2] = int(fftpack.next_fast_len(fast[-2] + 1))
    return fast


class Fourier:
    """A real image together with a cache of its (padded, origin-shifted) real FFTs."""

    def __init__(self, image, image_fft=None):
        self._image = np.asarray(image)
        self._fft = {} if image_fft is None else image_fft

    @property
    def image(self):
        return self._image

    @property
    def shape(self):
        return self._image.shape

    def __len__(self):
        return len(self._image)

    def __getitem__(self, i):
        return Fourier(self._image[i])

    def fft(self, fft_shape, axes):
        key = (tuple(int(s) for s in fft_shape), tuple(np.atleast_1d(axes).tolist()))
        if key not in self._fft:
            ax = key[1]
            padded = _pad(np.asarray(self._image, dtype=np.float64), key[0], ax)
            self._fft[key] = np.fft.rfftn(np.fft.ifftshift(padded, ax), axes=ax)
        return self._fft[key]

    @staticmethod
    def from_fft(image_fft, fft_shape, image_shape, axes=None):
        if axes is None:
            axes = tuple(range(len(fft_shape)))
        img = np.fft.fftshift(np.fft.irfftn(image_fft, fft_shape, axes=axes), axes=axes)
        return Fourier(_centered(img, image_shape))


def _as_fourier(x):
    return x if isinstance(x, Fourier) else Fourier(x)


def match_psf(psf1, psf2, padding=3, axes=(-2, -1)):
    """Difference kernel K with ``psf1 = K * psf2`` (k-space division).  Output box: the operand with more
    entries along axis 0, ties -> ``psf1``."""
    psf1, psf2 = _as_fourier(psf1), _as_fourier(psf2)
    if psf1.image.ndim != psf2.image.ndim:
        raise ValueError("Both images must have the same number of axes")
    fshape = _get_fft_shape(psf1.image, psf2.image, padding, axes)
    ratio = psf1.fft(fshape, axes) / psf2.fft(fshape, axes)
    shape = psf2.shape if len(psf1) < len(psf2) else psf1.shape
    return Fourier.from_fft(ratio, fshape, shape, axes)


def kernel_transform(kernel, image_shape, padding=3):
    """K^ of a (C, P, P) kernel on the fast grid of a (C, Ny, Nx) image -> (fft_shape, complex128 array)."""
    kernel = np.asarray(kernel, dtype=np.float64)
    fshape = _get_fft_shape(image_shape, kernel.shape, padding, (1, 2))
    return fshape, np.ascontiguousarray(_as_fourier(kernel).fft(fshape, (1, 2)))


def device_grid(image_shape, kernel_shape, padding=3):
    """FFT grid of the device fitting loop for a (C, Ny, Nx) frame and a (C, Py, Px) kernel.

    -> ``(Fy, Fx), (y0, x0)``.  ``(y0, x0)`` is the grid index of kernel pixel (0, 0) before wrapping under the
    reference's rule (centre-pad to ITS fast shape, then ``ifftshift``; fft.py:82-113, 255-273) -- ``-(P//2)`` for
    odd P.  The device keeps the frame at the grid origin and only ever reads back ``[0,N)``, so the grid only has
    to be long enough that nothing wraps INTO the frame: ``F >= N + max(-k0, P-1+k0)``, which is shorter than the
    reference's ``N + P + 3`` (e.g. 288 instead of 300 for N=256, P=41).  The convolution result inside the frame is
    the same linear convolution; only the rounding differs.  Lengths are taken from the set the fused spectral kernels
    are instantiated for."""
    ref = _get_fft_shape(image_shape, kernel_shape, padding, (1, 2))
    shape, origin = [], []
    for F, N, P in zip(ref, image_shape[1:], kernel_shape[1:]):
        k0 = (F - P + 1) // 2 - F // 2
        need = int(N + max(-k0, P - 1 + k0))
        f = int(nat.lib().sb_fft_supported_length(need))  # lengths of the fused spectral kernels (csrc/spectral.cuh)
        if f == 0:  # larger than any of them: any even 5-smooth length will do for cuFFT
            f = int(fftpack.next_fast_len(need))
            while f % 2:
                f = int(fftpack.next_fast_len(f + 1))
        if os.environ.get("SB_REFERENCE_GRID"):  # diagnostic: the reference's own (longer) fast shape
            f = int(F)
        shape.append(f)
        origin.append(int(k0))
    return tuple(shape), tuple(origin)


def device_convolve(image, khat, fshape, adjoint=False):
    """(C, Ny, Nx) image x precomputed K^ on the GPU; float32 in -> float32 out, anything else -> float64."""
    image = np.asarray(image)
    if image.ndim != 3:
        raise ValueError("device convolution expects a (C, Ny, Nx) cube")
    dt = np.float32 if image.dtype == np.float32 else np.float64
    img = nat.as_array(image, dt)
    out = np.empty_like(img)
    kh = np.ascontiguousarray(np.broadcast_to(khat, (img.shape[0],) + khat.shape[1:]), dtype=np.complex128)
    fn = nat.lib().sb_fft_convolve_f32 if dt == np.float32 else nat.lib().sb_fft_convolve_f64
    nat.check(fn(nat.ptr(img), img.shape[0], img.shape[1], img.shape[2], nat.ptr(kh), int(fshape[0]), int(fshape[1]),
                 int(bool(adjoint)), nat.ptr(out), nat.default_device()))
    return out


def convolve(image, kernel, padding=3, axes=(-2, -1), return_Fourier=True):
    """FFT convolution of a cube with a per-band kernel over the last two axes, cropped to the image shape."""
    image, kernel = _as_fourier(image), _as_fourier(kernel)
    img, ker = image.image, kernel.image
    if img.ndim == 2:
        img, ker = img[None], ker[None]
    if img.ndim != 3 or tuple(a % img.ndim for a in axes) != (1, 2):
        raise NotImplementedError("the device convolution covers (C, Ny, Nx) cubes over the spatial axes")
    fshape, khat = kernel_transform(ker, img.shape, padding)
    out = device_convolve(img, khat, fshape)
    if image.image.ndim == 2:
        out = out[0]
    return Fourier(out) if return_Fourier else out


def shift(image, shift, fft_shape=None, axes=(-2, -1), return_Fourier=True):
    """Sub-pixel translation by a Fourier phase ramp (scarlet/fft.py:399-428): centre-pad to the fast shape of
    (image, image, padding 10), ifftshift, rfftn, multiply by exp(-2 pi i (fftfreq_y s0 + rfftfreq_x s1)), irfftn, fftshift,
    centre-crop.  Host NumPy, for stand-alone ``get_model`` calls; inside the fitting loop the same linear map is evaluated on
    the device from its Toeplitz form (csrc/kernels.cuh: k_shift_apply)."""
    img = image.image if isinstance(image, Fourier) else np.asarray(image)
    if img.ndim != 2 or tuple(a % 2 for a in axes) != (0, 1):
        raise NotImplementedError("shift covers 2-D images")
    if fft_shape is None:
        fft_shape = _get_fft_shape(img, img, padding=10, axes=(0, 1))
    spec = _as_fourier(np.asarray(img, dtype=np.float64)).fft(fft_shape, (0, 1))
    ramp = np.exp(-2j * np.pi * np.fft.fftfreq(fft_shape[0]) * shift[0])[:, None] * \
        np.exp(-2j * np.pi * np.fft.rfftfreq(fft_shape[1]) * shift[1])[None, :]
    out = Fourier.from_fft(spec * ramp, fft_shape, img.shape, (0, 1))
    return out if return_Fourier else np.real(out.image)
