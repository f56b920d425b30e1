"""WCS helpers and sinc interpolation for multi-resolution scenes (host setup only).

Mirrors the slice of scarlet/interpolation.py that ``Frame.from_observations`` and ``ResolutionRenderer`` use at
set-up time: ``get_affine`` 378-384, ``get_pixel_size`` 387-394, ``get_angles`` 397-424, ``sinc_interp`` 427-503 (aligned
grids only), ``sinc_interp_inplace`` 505-560, ``get_psf_size`` 708-740.  Nothing here runs per iteration.
"""
import numpy as np

from . import fft


def get_affine(wcs):
    try:
        return wcs.wcs.pc
    except AttributeError:
        return wcs.cd


def get_pixel_size(affine):
    return np.sqrt(np.abs(affine[0, 0]) * np.abs(affine[1, 1] - affine[0, 1] * affine[1, 0]))


def get_angles(frame_wcs, model_wcs):
    """-> ([cos, sin] of the rotation between the two pixel grids, pixel-scale ratio frame/model)."""
    model_affine, frame_affine = np.asarray(get_affine(model_wcs)), np.asarray(get_affine(frame_wcs))
    model_pix, frame_pix = get_pixel_size(model_affine), get_pixel_size(frame_affine)
    h = frame_pix / model_pix
    fv = np.sum(frame_affine, axis=0)[:2] / frame_pix
    mv = np.sum(model_affine, axis=0)[:2] / model_pix
    fv = fv / np.sum(fv ** 2) ** 0.5
    mv = mv / np.sum(mv ** 2) ** 0.5
    sin_rot = fv[0] * mv[1] - fv[1] * mv[0]
    cos_rot = np.dot(fv, mv)
    return [cos_rot, sin_rot], h


def sinc_interp(images, coord_hr, coord_lr, angle=None, padding=3):
    """Whittaker-Shannon interpolation of a cube sampled at ``coord_lr`` onto ``coord_hr``.

    Two branches, as in the reference (interpolation.py:427-503): separable sinc matrices when ``1 - cos < eps``, else
    Fourier shifts along the rotated axis followed by sinc sums.  (An identity rotation computed from an affine matrix
    typically has ``cos = 1 - 2.2e-16``, which is NOT below eps, so aligned grids usually take the second branch.)"""
    y_hr, x_hr = coord_hr
    y_lr, x_lr = coord_lr
    hy, hx = np.abs(y_lr[1] - y_lr[0]), np.abs(x_lr[1] - x_lr[0])
    assert hy != 0 and hx != 0
    if angle is None or 1 - angle[0] < np.finfo(float).eps:
        sy = np.sinc((y_lr[np.newaxis, :] - y_hr[:, np.newaxis]) / hy)
        sx = np.sinc((x_lr[:, np.newaxis] - x_hr[np.newaxis, :]) / hx)
        return np.array([np.dot(np.dot(sy, image.T), sx) for image in images])
    cos, sin = angle[0], angle[1]
    images = np.asarray(images)
    fshape = fft._get_fft_shape(images, images, padding=padding, axes=[1, 2])
    x_fft = fft.Fourier(images).fft(fshape, (1, 2))
    shifter_y = -2j * np.pi * np.fft.fftfreq(fshape[0])
    shifter_x = -2j * np.pi * np.fft.rfftfreq(fshape[1])
    shift_y = np.exp(shifter_y[np.newaxis, :] * (-(y_hr[:, np.newaxis]) * cos))
    shift_x = np.exp(shifter_x[np.newaxis, :] * (-(y_hr[:, np.newaxis]) * sin))
    res_fft = x_fft[:, np.newaxis, :, :] * shift_y[np.newaxis, :, :, np.newaxis] * shift_x[np.newaxis, :, np.newaxis, :]
    shape = (res_fft.shape[0], res_fft.shape[1], images.shape[1], images.shape[2])
    shifted = fft.Fourier.from_fft(res_fft, fshape, shape, [2, 3]).image
    shy = np.sinc((y_lr[np.newaxis, :] + x_hr[:, np.newaxis] * sin) / hy)
    shx = np.sinc((x_lr[np.newaxis, :] - x_hr[:, np.newaxis] * cos) / hx)
    # out[c, i, a] = sum_x (sum_y shifted[c, i, y, x] shy[a, y]) shx[a, x].  The y sum runs row by row in index order -- the
    # order of the reference's reduction over that axis; the deconvolved difference kernel amplifies a different summation
    # order (a matrix product) to 1e-10 -- without materialising the (C, n_y_hr, n_x_hr, Ny, Nx) product.
    # Blocks of output rows keep the accumulator in cache.
    n_c, n_i, n_y, n_x = shifted.shape
    out = np.empty((n_c, n_i, shy.shape[0]), dtype=np.result_type(shifted, shy))
    rows = max(1, (1 << 18) // max(1, n_c * shy.shape[0] * n_x))
    for i0 in range(0, n_i, rows):
        blk = shifted[:, i0:i0 + rows]
        acc = np.zeros((n_c, blk.shape[1], shy.shape[0], n_x), dtype=out.dtype)
        tmp = np.empty_like(acc)
        for y in range(n_y):
            np.multiply(blk[:, :, np.newaxis, y, :], shy[np.newaxis, np.newaxis, :, y, np.newaxis], out=tmp)
            acc += tmp
        out[:, i0:i0 + rows] = (acc * shx[np.newaxis, np.newaxis, :, :]).sum(axis=-1)
    return out


def sinc_interp_inplace(image, h_image, h_target, angle, pad_shape=None):
    """Resample a cube from pixel scale ``h_image`` to ``h_target`` over the same physical area (odd output size)."""
    assert image.ndim == 3
    if pad_shape is not None:
        image = fft._pad(image, pad_shape, axes=[-2, -1])
    ny_lr, nx_lr = image.shape[-2:]
    coord_lr = np.array([np.arange(ny_lr) - (ny_lr - 1) / 2, np.arange(nx_lr) - (nx_lr - 1) / 2])
    ny_hr = int(np.round(ny_lr * h_image / h_target))
    nx_hr = int(np.round(nx_lr * h_image / h_target))
    ny_hr += ny_hr % 2 == 0
    nx_hr += nx_hr % 2 == 0
    coord_hr = np.array([np.arange(ny_hr) - (ny_hr - 1) / 2, np.arange(nx_hr) - (nx_hr - 1) / 2]) / h_image * h_target
    return sinc_interp(image, coord_hr, coord_lr, angle=angle)


def get_psf_size(psf):
    """3-sigma radius (pixels) estimated from the area above half maximum."""
    frame = psf / np.max(psf)
    area = np.sum(frame > 0.5)
    d = 2 * (area / np.pi) ** 0.5
    return 3 * d / (2 * (2 * np.log(2)) ** 0.5)


def shift_weights(F, shifts):
    """(len(shifts), F) complex matrix ``exp(-2 pi i f_k s)`` of a Fourier shift by ``s`` pixels on an F-periodic grid
    with the real-transform semantics of the reference (``mk_shifter(real=True)`` + ``irfftn``, interpolation.py:341-375,
    renderer.py:414-476): signed frequencies, and the Nyquist bin of an even grid keeps only its real part, cos(pi s)."""
    s = np.asarray(shifts, dtype=np.float64)
    k = np.arange(F)
    f = np.where(k <= F // 2, k, k - F) / F
    W = np.exp(-2j * np.pi * f[np.newaxis, :] * s[:, np.newaxis])
    if F % 2 == 0:
        W[:, F // 2] = np.cos(np.pi * s)
    return W


def shift_multiplier(Fy, Fx, s0, s1):
    """Fourier multiplier, on the half plane ``ky in [0, Fy), kx in [0, Fx/2]``, of the reference's two-axis Fourier shift by
    ``(s0, s1)`` pixels: ``rfftn`` over (y, x), phase ramp ``exp(-2 pi i (fftfreq_y s0 + rfftfreq_x s1))``, ``irfftn`` back
    (renderer.py:414-476 with axes (1, 2), ``mk_shifter(real=False)``).  The real inverse transform keeps only the Hermitian
    part of what the ramp produces, which matters on the Nyquist lines of even grids; in real space the operation is
    ``Re(Cy) u Re(Tx)^T - Im(Cy) u Im(Tx)^T`` with the circulant matrices of DESIGN 3.8, whose symbols give

        sigma(ky, kx) = ReCy(ky) ReTx(kx) - ImCy(ky) ImTx(kx).

    ``s0``, ``s1`` may be arrays of n shifts -> (n, Fy, Fx/2+1) complex128.  Checked against the reference's own rotated render
    to 4e-14 (tests/golden/multires_rot.npz)."""
    s0 = np.atleast_1d(np.asarray(s0, dtype=np.float64))[:, None]
    s1 = np.atleast_1d(np.asarray(s1, dtype=np.float64))[:, None]
    ky = np.arange(Fy)
    my = np.where(ky < (Fy + 1) // 2, ky, ky - Fy)  # numpy.fft.fftfreq: the Nyquist bin of an even grid is negative
    cy = np.exp(-2j * np.pi * my[None, :] * s0 / Fy)
    cym = cy[:, (-ky) % Fy]
    re_cy, im_cy = (cy + np.conj(cym)) / 2, (cy - np.conj(cym)) / 2j
    kx = np.arange(Fx)
    ck = np.where((kx == 0) | (2 * kx == Fx), 1.0, 2.0)
    tx = np.where(kx[None, :] <= Fx // 2, ck[None, :] * np.exp(-2j * np.pi * kx[None, :] * s1 / Fx), 0)
    txm = tx[:, (-kx) % Fx]
    re_tx, im_tx = ((tx + np.conj(txm)) / 2)[:, :Fx // 2 + 1], ((tx - np.conj(txm)) / 2j)[:, :Fx // 2 + 1]
    return re_cy[:, :, None] * re_tx[:, None, :] - im_cy[:, :, None] * im_tx[:, None, :]
