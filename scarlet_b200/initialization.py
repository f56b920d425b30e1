"""Data-driven initialisation of sources: the step right BEFORE the fitting path.  Host code, once per
source; the only heavy operator it uses, the radial-monotonicity sweep over the detection image, runs on the GPU.

Mirrors the slice of scarlet/initialization.py behind ``ExtendedSource`` / ``PointSource``: ``get_pixel_spectrum`` 12-85,
``get_minimal_boxsize`` 173-177, ``trim_morphology`` 180-210, ``build_initialization_image`` 213-284, and the morphology
recipes of scarlet/source.py (``CompactExtendedSource.init_morph`` 316-363, ``SingleExtendedSource.init_morph`` 453-522).
"""
import logging

import numpy as np

from . import operator
from .bbox import Box, overlapped_slices
from .constraint import CenterOnConstraint
from .morphology import get_minimal_boxsize  # noqa: F401  (re-exported under the reference's module name)
from .renderer import ConvolutionRenderer, NullRenderer

logger = logging.getLogger("scarlet_b200.initialization")


def _as_tuple(observations):
    return tuple(observations) if hasattr(observations, "__iter__") else (observations,)


def get_pixel_spectrum(sky_coord, observations, correct_psf=False, models=None, concat=True):
    """Data values at the pixel nearest to ``sky_coord`` in every observation (optionally divided by the PSF peak, or by
    the value of a rendered unit-flux model there): the spectrum of a compact source sitting on that pixel."""
    single = not hasattr(observations, "__iter__")
    observations = _as_tuple(observations)
    if models is None:
        models = (None,) * len(observations)
    elif single:
        models = (models,)
    assert len(models) == len(observations)
    assert not (correct_psf and any(m is not None for m in models))
    spectra = []
    for obs, model in zip(observations, models):
        iy, ix = np.round(obs.get_pixel(sky_coord)).astype(int)
        spectrum = np.array(obs.data[:, iy, ix])
        if correct_psf and obs.psf is not None:
            spectrum = spectrum / obs.psf.get_model().max(axis=(1, 2))
        elif model is not None:
            spectrum = spectrum / np.asarray(model)[:, iy, ix]
        if np.any(spectrum <= 0):
            (logger.warning if np.all(spectrum <= 0) else logger.info)("Zero or negative spectrum %s at %s", spectrum, sky_coord)
        spectra.append(spectrum)
    return np.concatenate(spectra).reshape(-1) if concat else spectra


def build_initialization_image(observations, spectra=None):
    """Inverse-variance, spectrum-weighted coadd of all same-grid observations in the model frame and its noise level:
    the detection image a source morphology is cut from."""
    single = not hasattr(observations, "__iter__")
    observations = _as_tuple(observations)
    if spectra is None or single:
        spectra = (spectra,) if single else (None,) * len(observations)
    assert len(spectra) == len(observations)
    frame = observations[0].model_frame
    usable = [i for i, obs in enumerate(observations) if isinstance(obs.renderer, (NullRenderer, ConvolutionRenderer))]
    cached = getattr(observations[0], "_detect", None)
    if cached is None:
        detect = np.zeros((len(usable),) + tuple(frame.shape), dtype=frame.dtype)
        var = np.zeros_like(detect)
        for n, i in enumerate(usable):
            obs = observations[i]
            data_sl, model_sl = obs.renderer.slices
            obs.renderer.map_channels(detect[n])[model_sl] += obs.data[data_sl]
            obs.renderer.map_channels(var[n])[model_sl] += np.ma.filled(obs.noise_rms, np.inf)[data_sl] ** 2
        var[~np.isfinite(var)] = 0  # masked pixels carry no weight
        cached = observations[0]._detect = (detect, var)
    detect, var = cached
    sed = np.zeros((len(usable), frame.C))
    for n, i in enumerate(usable):
        observations[i].renderer.map_channels(sed[n])[:] = 1 if spectra[i] is None else spectra[i]
    sed = sed[:, :, None, None]
    weight = np.zeros(var.shape)
    np.divide(1, var, out=weight, where=var > 0)
    weight = weight * sed
    return (weight * detect).sum(axis=(0, 1)), np.sqrt((sed * weight).sum(axis=(0, 1)))


def trim_morphology(center_index, morph, bg_thresh=0, boxsize=None):
    """Zero everything at or below ``bg_thresh`` and cut the smallest allowed square box around ``center_index`` that holds
    what is left.  -> (morph in that box, 2-D Box)"""
    morph[morph <= bg_thresh] = 0
    support = Box.from_data(morph, min_value=0)
    size = 0
    if support.contains(center_index):
        size = 2 * max(center_index[0] - support.start[-2], support.stop[0] - center_index[-2],
                       center_index[1] - support.start[-1], support.stop[1] - center_index[-1])
    if boxsize is None:
        boxsize = get_minimal_boxsize(size)
    half = boxsize // 2
    bbox = Box.from_bounds((center_index[0] - half, center_index[0] + half + 1), (center_index[1] - half, center_index[1] + half + 1))
    return bbox.extract_from(morph), bbox


def compact_morphology(frame, sky_coord, boxsize=None):
    """Band-averaged model PSF centred on the source pixel, peak-normalised, in a standard box (source.py:316-363)."""
    center_index = np.round(frame.get_pixel(sky_coord)).astype(int)
    psf = frame.psf.get_model().mean(axis=0)
    psf_box = Box(psf.shape, origin=(center_index[0] - psf.shape[0] // 2, center_index[1] - psf.shape[1] // 2))
    if boxsize is None:
        boxsize = get_minimal_boxsize(max(psf.shape))
    bbox = Box((boxsize, boxsize), origin=(center_index[0] - boxsize // 2, center_index[1] - boxsize // 2))
    morph = np.zeros((boxsize, boxsize))
    into, frm = overlapped_slices(bbox, psf_box)
    morph[into] = psf[frm]
    return morph / morph.max(), bbox


def extended_morphology(frame, sky_coord, detect, detect_std, thresh=1, symmetric=True, monotonic="flat", min_grad=0, boxsize=None):
    """Symmetrised (minimum of partner pixels), radially monotonic cut-out of the detection image above
    ``thresh * detect_std``, peak-normalised and floored by the compact (PSF) morphology (source.py:453-522)."""
    center_index = np.round(frame.get_pixel(sky_coord)).astype(int)
    im = np.array(detect, dtype=np.float64)
    if symmetric:
        im = operator.prox_uncentered_symmetry(im, 0, center=center_index, algorithm="sdss")
    if monotonic:
        kind = "angle" if monotonic is True else monotonic
        im = operator.windowed_monotonic(im, center_index, neighbor_weight=kind, min_gradient=min_grad)
    morph, bbox = trim_morphology(center_index, im, bg_thresh=detect_std * thresh, boxsize=boxsize)
    if morph.sum() > 0:
        morph = morph / morph.max()
    else:
        logger.warning("No flux in morphology model for source at %s", sky_coord)
        morph = CenterOnConstraint(tiny=1)(morph, 0)
    if frame.psf is not None:
        psf_morph, _ = compact_morphology(frame, sky_coord, boxsize=max(bbox.shape))
        morph = np.maximum(morph, psf_morph)
    return morph, bbox
