"""Model frame -> observation frame mappings.  Mirrors scarlet/renderer.py: ``Renderer`` 13-83,
``NullRenderer`` 86-94, ``match_shape`` 130-161, ``ConvolutionRenderer`` 164-259.

Setup (channel map, data/model overlap, difference kernel, K^) happens here on the host; the convolution itself
runs on the device: inside the fitting loop as part of the plan, and for a stand-alone ``renderer(model)`` call
through ``sb_fft_convolve_*``.
"""
import numpy as np

from . import fft
from .bbox import Box, overlapped_slices
from .model import Model


class Renderer(Model):
    def __init__(self, data_frame, model_frame, *parameters):
        self.data_frame = data_frame
        self.model_frame = model_frame
        self.channel_map = self.get_channel_map(data_frame, model_frame)
        super().__init__(*parameters)

    def __call__(self, model, *parameters):
        return self.get_model(*parameters)(model)

    def get_channel_map(self, data_frame, model_frame):
        if list(data_frame.channels) == list(model_frame.channels):
            return None
        idx = [list(model_frame.channels).index(c) for c in data_frame.channels]
        lo, hi = min(idx), max(idx)
        if hi + 1 - lo == len(idx) and idx == list(range(lo, hi + 1)):
            return slice(lo, hi + 1)
        raise NotImplementedError("non-contiguous channel maps are not supported (the reference returns the index "
                                  "list, which its own map_channels cannot apply either)")

    def map_channels(self, model):
        if self.channel_map is None:
            return model
        return model[self.channel_map]

    @property
    def channel_offset(self):
        return 0 if self.channel_map is None else self.channel_map.start


def _spatial_slices(data_frame, model_frame):
    """Overlap of the data pixels with the model frame (translation only)."""
    pix = np.atleast_2d(data_frame.convert_pixel_to(model_frame))
    ll = np.round(pix.min(axis=0)).astype(int)
    ur = np.round(pix.max(axis=0)).astype(int) + 1
    data_box = model_frame.bbox[0] @ Box.from_bounds((ll[0], ur[0]), (ll[1], ur[1]))
    return overlapped_slices(data_box, model_frame.bbox), (int(ll[0]), int(ll[1]))


def match_shape(model, data_frame, slices):
    data_slices, model_slices = slices
    if any(data_slices[d].stop - data_slices[d].start != data_frame.shape[d] for d in (-2, -1)):
        out = np.zeros(data_frame.shape, dtype=data_frame.dtype)
        out[data_slices] = model[model_slices]
        return out
    return model[model_slices]


class NullRenderer(Renderer):
    """Observation and model share the PSF: channel selection + shape matching only."""

    def __init__(self, data_frame, model_frame):
        super().__init__(data_frame, model_frame)
        self.slices, self.origin = _spatial_slices(data_frame, model_frame)
        self.diff_kernel = None

    def get_model(self, *parameters):
        return lambda model: match_shape(self.map_channels(model), self.data_frame, self.slices)


class ConvolutionRenderer(Renderer):
    def __init__(self, data_frame, model_frame, *parameters, convolution_type="fft", padding=10, psf_shift=None):
        if psf_shift is not None:
            # a fitted offset of the PSF difference kernel (astrometric mismatch between observations): renderer parameter
            # with step 1e-2 and no constraint (renderer.py:175-177); it joins the optimiser's parameter tuple behind the
            # sources' parameters (blend.py:103-105).  On the device: csrc/psf_shift.cuh.
            from .parameter import Parameter
            if not isinstance(psf_shift, Parameter):
                psf_shift = Parameter(np.asarray(psf_shift, dtype=np.float64), name="psf_shift", step=1.0e-2)
            elif psf_shift.name != "psf_shift":
                raise AssertionError("the renderer parameter must be named 'psf_shift'")
            parameters = (*parameters, psf_shift)
        if convolution_type not in ("fft", "real"):
            raise ValueError("`convolution` must be either 'real' or 'fft', got {}".format(convolution_type))
        # "real" (renderer.py:97-127, operators_pybind11.cc:39-56: a sum of shifted, scaled copies of the image, zero outside
        # the frame) and "fft" are the same linear map -- a "same"-size convolution with the difference kernel -- and the
        # device evaluates that one map for both; only the floating-point summation order differs from the reference's
        # real-space loop.
        super().__init__(data_frame, model_frame, *parameters)
        self._convolution_type = convolution_type
        self.slices, self.origin = _spatial_slices(data_frame, model_frame)
        psf_obs = fft.Fourier(data_frame.psf.get_model().astype(model_frame.dtype))
        psf_model = fft.Fourier(model_frame.psf.get_model().astype(model_frame.dtype))
        self.diff_kernel = fft.match_psf(psf_obs, psf_model, padding=padding)
        self._khat = None

    def kernel_transform(self):
        """(fft_shape, K^) of the difference kernel on the grid of the model-frame sub-cube."""
        if self._khat is None:
            sub_shape = (self.data_frame.C,) + tuple(self.model_frame.shape[1:])
            self._khat = fft.kernel_transform(self.diff_kernel.image, sub_shape, padding=3)
        return self._khat

    def device_kernel(self):
        """(fft_shape, (y0, x0), float64 kernel image) for the device fitting loop, which transforms the kernel itself."""
        sub_shape = (self.data_frame.C,) + tuple(self.model_frame.shape[1:])
        ker = np.ascontiguousarray(self.diff_kernel.image, dtype=np.float64)
        fshape, origin = fft.device_grid(sub_shape, ker.shape, padding=3)
        return fshape, origin, ker

    def shifted_kernel(self, psf_shift):
        """The difference kernel moved by ``psf_shift`` (renderer.py:220-227: ``fft.shift`` per band on the fast grid of
        (kernel, kernel, padding 10), cropped back to the kernel box)."""
        return np.stack([fft.shift(k, psf_shift, return_Fourier=False) for k in np.asarray(self.diff_kernel.image, dtype=np.float64)])

    def shift_grid(self):
        """fast grid of ``fft.shift`` for the kernel image"""
        ker = np.asarray(self.diff_kernel.image)
        return tuple(int(f) for f in fft._get_fft_shape(ker[0], ker[0], padding=10, axes=(0, 1)))

    def convolve(self, model, convolution_type=None, psf_shift=None):
        if psf_shift is not None:
            sub_shape = (self.data_frame.C,) + tuple(self.model_frame.shape[1:])
            fshape, khat = fft.kernel_transform(self.shifted_kernel(np.asarray(psf_shift)), sub_shape, padding=3)
        else:
            fshape, khat = self.kernel_transform()
        return fft.device_convolve(np.asarray(model), khat, fshape)

    def get_model(self, *parameters):
        shift = self.get_parameter("psf_shift", *parameters)

        def transform(model, *ignored):
            return match_shape(self.convolve(self.map_channels(model), psf_shift=shift), self.data_frame, self.slices)
        return transform


class ResolutionRenderer(Renderer):
    """Observation on a different pixel grid than the model: resample + convolve (scarlet/renderer.py:262-547).

    Set-up follows the reference: difference kernel between the observed PSF, sinc-resampled to the model pixel scale,
    and the model PSF (``build_diffkernel`` 365-412); fast grid ``_fft_shape``; positions of the low-resolution pixel rows
    and columns in model-frame coordinates (``shifts`` 308-316), for aligned and for rotated grids (318-347).

    The reference then tabulates, for every low-resolution row, the kernel Fourier-shifted to that row
    (``_resconv_op``, (C, n_y, Fy*Fx)) and evaluates a render as Fourier shifts of the model to every low-resolution
    column followed by a skinny matrix product (478-547).  Both steps are shifts on the same periodic grid, so with
    Parseval's theorem the render is

        LR[c,i,j] = h^2/(Fy Fx) * sum_{ky,kx} Ey[i,ky] Ex[j,kx] K^[c,ky,kx] conj(M^[c,ky,kx])

    with ``Ey/Ex = exp(-2 pi i f s)`` (Nyquist bins real, the ``irfftn`` semantics of the reference).  This identity is
    what the device evaluates (csrc/spectral.cuh: resampling kernels) and what ``get_model`` evaluates in NumPy for
    stand-alone ``Observation.render`` calls; it agrees with the reference's own render to 1e-13 in float64
    (tests/golden/multires.npz).
    """

    def __init__(self, data_frame, model_frame, padding=10):
        from . import interpolation
        super().__init__(data_frame, model_frame)
        self.angle, self.h = interpolation.get_angles(data_frame.wcs, model_frame.wcs)
        self.isrot = (np.abs(self.angle[1]) ** 2) > np.finfo(float).eps
        lr_shape = data_frame.shape[1:]
        if lr_shape[0] != lr_shape[1]:
            raise ValueError("ResolutionRenderer needs square observations (as the reference does, renderer.py:274)")
        pixels = np.stack((np.arange(lr_shape[0]), np.arange(lr_shape[1])), axis=1)
        coord_hr = np.array(data_frame.convert_pixel_to(model_frame, pixel=pixels), dtype=np.float64)
        diff_psf, psf_hr = self.build_diffkernel(data_frame, model_frame)
        self.small_axis = data_frame.Nx <= data_frame.Ny
        self._fft_shape = fft._get_fft_shape(psf_hr, np.zeros(model_frame.shape), padding=3, axes=[-2, -1], max=False)
        if self._fft_shape[-2] < diff_psf.shape[-2] or self._fft_shape[-1] < diff_psf.shape[-1]:
            diff_psf = fft.Fourier(fft._centered(diff_psf.image, np.array([diff_psf.shape[0] + 1, *self._fft_shape]) - 1))
        self.diff_kernel = fft.Fourier(fft._pad(diff_psf.image, self._fft_shape, axes=(-2, -1)))
        Fy, Fx = self._fft_shape
        center_y = int(Fy / 2.0 - (Fy - model_frame.Ny) / 2.0) + ((Fy % 2) != 0) * ((model_frame.Ny % 2) == 0)
        center_x = int(Fx / 2.0 - (Fx - model_frame.Nx) / 2.0) - ((Fx % 2) != 0) * ((model_frame.Nx % 2) == 0)
        if not self.isrot:
            self.shifts = coord_hr.T.copy()
            self.shifts[0] -= center_y
            self.shifts[1] -= center_x
            self.other_shifts = np.copy(self.shifts)
        else:
            # rotated grids (renderer.py:318-347): positions of the low-resolution rows / columns along the observation's own
            # axes; a row then needs a shift along both model axes, (Y cos, -Y sin), a column (X sin, X cos)
            cos, sin = float(self.angle[0]), float(self.angle[1])
            self.Y_unrot = ((coord_hr[:, 0] - center_y) * cos - (coord_hr[:, 1] - center_x) * sin).reshape(lr_shape[0])
            self.X_unrot = ((coord_hr[:, 1] - center_x) * cos + (coord_hr[:, 0] - center_y) * sin).reshape(lr_shape[1])
            rows = np.array([self.Y_unrot * cos, -self.Y_unrot * sin])
            cols = np.array([sin * self.X_unrot, cos * self.X_unrot])
            self.shifts, self.other_shifts = (rows, cols) if self.small_axis else (cols, rows)
        self.origin = (0, 0)
        self._operator = None

    def build_diffkernel(self, data_frame, model_frame):
        from . import interpolation
        psf_hr = np.array(model_frame.psf.get_model(), dtype=np.float64)
        psf_lr = np.array(data_frame.psf.get_model()).astype(model_frame.dtype)
        pad_shape = np.array((np.array(data_frame.shape[-2:]) + np.array(psf_lr.shape[-2:])) / 2).astype(int) * 2 + 1
        h_lr = interpolation.get_pixel_size(np.asarray(interpolation.get_affine(data_frame.wcs)))
        h_hr = interpolation.get_pixel_size(np.asarray(interpolation.get_affine(model_frame.wcs)))
        angle, _ = interpolation.get_angles(model_frame.wcs, data_frame.wcs)
        psf_lr_hr = interpolation.sinc_interp_inplace(psf_lr, h_lr, h_hr, angle, pad_shape=pad_shape)
        psf_hr = psf_hr / np.sum(psf_hr)
        psf_lr_hr = psf_lr_hr / np.sum(psf_lr_hr)
        return fft.match_psf(fft.Fourier(psf_lr_hr), fft.Fourier(psf_hr)), psf_hr

    def device_operator(self):
        """What the device needs: K^ of the centred-padded kernel (half spectrum along x, complex128 (C, Fy, Fx/2+1)), the
        shift matrices for the model frame stored at the grid origin, and the overall scale h^2 / (Fy Fx)."""
        if self._operator is None:
            from . import interpolation
            Fy, Fx = (int(f) for f in self._fft_shape)
            Ny, Nx = self.model_frame.shape[1:]
            oy, ox = (Fy - Ny + 1) // 2, (Fx - Nx + 1) // 2  # where fft._pad puts the model inside the grid
            khat = np.ascontiguousarray(np.fft.rfft2(np.asarray(self.diff_kernel.image, dtype=np.float64), axes=(1, 2)))
            if self.isrot:
                # LR[c,i,j] = scale sum_kx w_kx Re sum_ky K^ conj(M^) A_i B_j: A_i / B_j are the half-plane multipliers of the
                # two-axis shifts attached to row i / column j.  The kernel is shifted by `shifts`, the model (centred in the
                # grid by the reference, at the grid origin on the device: an integer translation by (oy, ox)) by
                # -`other_shifts`; the model's multiplier enters conjugated.
                f_y = np.fft.fftfreq(Fy)[:, None]
                f_x = np.fft.rfftfreq(Fx)[None, :]
                kernel_mult = interpolation.shift_multiplier(Fy, Fx, self.shifts[0], self.shifts[1])
                model_mult = np.conj(interpolation.shift_multiplier(Fy, Fx, -self.other_shifts[0], -self.other_shifts[1])) * \
                    np.exp(2j * np.pi * (f_y * oy + f_x * ox))[None]
                A, B = (kernel_mult, model_mult) if self.small_axis else (model_mult, kernel_mult)
                self._operator = dict(fshape=(Fy, Fx), khat=khat, A=np.ascontiguousarray(A), B=np.ascontiguousarray(B),
                                      scale=float(self.h ** 2 / (Fy * Fx)), rotated=True)
                return self._operator
            Ey = np.ascontiguousarray(interpolation.shift_weights(Fy, self.shifts[0] - oy))
            Ex = np.ascontiguousarray(interpolation.shift_weights(Fx, self.shifts[1] - ox)[:, :Fx // 2 + 1])
            self._operator = dict(fshape=(Fy, Fx), khat=khat, Ey=Ey, Ex=Ex, scale=float(self.h ** 2 / (Fy * Fx)), rotated=False)
        return self._operator

    def get_model(self, *parameters):
        def transform(model):
            """Stand-alone render (host NumPy, float64): the forward identity of the class docstring."""
            op = self.device_operator()
            Fy, Fx = op["fshape"]
            m = np.asarray(self.map_channels(model), dtype=np.float64)
            mhat = np.fft.rfft2(m, s=(Fy, Fx), axes=(1, 2))
            wgt = np.full(Fx // 2 + 1, 2.0)
            wgt[0] = 1.0
            if Fx % 2 == 0:
                wgt[-1] = 1.0
            if op["rotated"]:
                p = op["khat"] * np.conj(mhat) * wgt
                nc, ni, nj = p.shape[0], op["A"].shape[0], op["B"].shape[0]
                u = (p[:, None] * op["A"][None]).reshape(nc * ni, -1)  # one (C H) x (Fy Fx/2+1) by (Fy Fx/2+1) x W product
                out = op["scale"] * (u @ op["B"].reshape(nj, -1).T).real.reshape(nc, ni, nj)
                return out.astype(np.asarray(model).dtype)
            t1 = np.einsum("iy,cyx->cix", op["Ey"], op["khat"] * np.conj(mhat))
            out = op["scale"] * np.einsum("cix,jx,x->cij", t1, op["Ex"], wgt).real
            return out.astype(np.asarray(model).dtype)
        return transform
