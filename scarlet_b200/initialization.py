"""Data-driven initialisation of sources: the step right BEFORE the fitting path.  Host code, once per
source; the only heavy operator it uses, the radial-monotonicity sweep over the detection image, runs on the GPU.

Mirrors the slice of scarlet/initialization.py behind ``ExtendedSource`` / ``PointSource``: ``get_pixel_spectrum`` 12-85,
``get_psf_spectrum`` 88-170, ``get_minimal_boxsize`` 173-177, ``trim_morphology`` 180-210, ``build_initialization_image``
213-284, ``init_all_sources`` 287-363, ``init_source`` 366-490, ``set_spectra_to_match`` 493-588, and the morphology
recipes of scarlet/source.py (``CompactExtendedSource.init_morph`` 316-363, ``SingleExtendedSource.init_morph`` 453-522).
"""
import logging

import numpy as np
import numpy.ma as ma

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
        # in-place divisions, like the reference's: the spectrum keeps the dtype of the data (the device rounds a float32
        # spectrum to float32 after every step, so the dtype is part of the result)
        if correct_psf and obs.psf is not None:
            spectrum /= obs.psf.get_model().max(axis=(1, 2))
        elif model is not None:
            spectrum /= np.asarray(model)[:, iy, ix]
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
            # The reference adds the MASKED noise array into a plain one, which takes the data underneath the mask: a
            # zero-weight pixel enters the coadd with unit variance (numpy.ma leaves the numerator of 1/sqrt(w) there).
            # Kept as it is: the initial morphologies are compared with the reference's.
            obs.renderer.map_channels(var[n])[model_sl] += np.asarray(ma.getdata(obs.noise_rms))[data_sl] ** 2
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
    morph[~(morph > bg_thresh)] = 0  # also clears NaN pixels
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


def get_psf_spectrum(sky_coord, observations, compute_snr=False, concat=True):
    """Point-source photometry at ``sky_coord``: in every channel the amplitude of the observation's PSF that best matches
    the data around that position (projection of the cut-out on the PSF image over the unmasked pixels), for all
    observations; with ``compute_snr`` also the matched-filter signal-to-noise ratio summed over all channels
    (initialization.py:88-170)."""
    observations = _as_tuple(observations)
    spectra, signal, variance = [], [], []
    for obs in observations:
        index = np.round(obs.get_pixel(sky_coord)).astype(int)
        psf = obs.psf.get_model()
        box = obs.psf.bbox + (0, *index)
        image = box.extract_from(obs.data)
        rms = obs.noise_rms
        noise = box.extract_from(np.asarray(ma.getdata(rms)))
        masked = box.extract_from(ma.getmaskarray(rms))  # pixels outside the observation count as unmasked zeros, like masked ones
        amplitudes = np.empty(obs.C, dtype=np.result_type(image.dtype, psf.dtype))
        for c in range(obs.C):
            keep = ~masked[c]
            p, d = psf[c][keep], image[c][keep]
            proj = d @ p
            amplitudes[c] = proj / (p @ p)
            if compute_snr:
                signal.append(proj)
                variance.append((p * noise[c][keep] ** 2) @ p)
        if np.any(amplitudes <= 0):
            (logger.warning if np.all(amplitudes <= 0) else logger.info)("Zero or negative spectrum %s at %s", amplitudes, sky_coord)
        spectra.append(amplitudes)
    if concat:
        spectra = np.concatenate(spectra).reshape(-1)
    if compute_snr:
        return spectra, np.sum(signal) / np.sqrt(np.sum(variance))
    return spectra


def init_source(frame, center, observations, thresh=1, max_components=1, min_components=1, min_snr=50, shifting=False,
                resizing=True, boxsize=None, fallback=True):
    """One source at ``center`` with as many ``ExtendedSource`` components as its signal-to-noise supports
    (initialization.py:366-490): with ``fallback`` the number of components is capped at ``floor(psf_snr / min_snr)`` (at least
    ``min_components``) and lowered by one whenever the model comes out non-finite, down to zero components = a compact
    (PSF-shaped) extended source; if that fails too, ``None`` is returned."""
    from .source import ExtendedSource
    observations = _as_tuple(observations)
    if fallback:
        _, psf_snr = get_psf_spectrum(center, observations, compute_snr=True)
        supported = np.floor(psf_snr / min_snr)  # components the signal-to-noise ratio can carry
        max_components = min(max_components, max(min_components, int(supported) if np.isfinite(supported) else min_components))
    while max_components >= 0:
        try:
            if max_components > 0:
                source = ExtendedSource(frame, center, observations, thresh=thresh, shifting=shifting, resizing=resizing,
                                        boxsize=boxsize, K=max_components)
            else:
                source = ExtendedSource(frame, center, observations, shifting=shifting, resizing=resizing, boxsize=boxsize,
                                        compact=True)
            source.check_parameters()  # ArithmeticError when a parameter is not finite
        except ArithmeticError as e:
            if not fallback:
                raise
            logger.info("Could not initialize source at %s with %d components: %s", center, max_components, e)
            max_components -= 1
            continue
        return source
    return None


def init_all_sources(frame, centers, observations, thresh=1, max_components=1, min_components=1, min_snr=50, shifting=False,
                     resizing=True, boxsize=None, fallback=True, silent=False, set_spectra=True):
    """``init_source`` for every entry of ``centers``; sources that raise are listed in ``skipped`` when ``silent`` (else the
    exception propagates); finally the spectra of all sources are solved for jointly (``set_spectra_to_match``).  Returns
    ``(sources, skipped)`` (initialization.py:287-363)."""
    observations = _as_tuple(observations)
    sources, skipped = [], []
    for k, center in enumerate(centers):
        try:
            sources.append(init_source(frame, center, observations, thresh=thresh, max_components=max_components,
                                       min_components=min_components, min_snr=min_snr, shifting=shifting, resizing=resizing,
                                       boxsize=boxsize, fallback=fallback))
        except Exception:
            logger.warning("Failed to initialize source %d", k)
            if not silent:
                raise
            skipped.append(k)
    if set_spectra:
        set_spectra_to_match(sources, observations)
    return sources, skipped


def set_spectra_to_match(sources, observations):
    """Best-fit spectra of all factorized components, given their morphologies (initialization.py:493-588): every component
    is rendered with a flat unit spectrum into every observation (``Observation.render``, on the device) and, channel by
    channel, the amplitudes solve the weighted linear least-squares problem ``(M W M^T) a = M W d``.  Components whose rendered
    flux falls mostly on zero-weight pixels of a channel are left out of that channel's solve (amplitude 0); components with
    identical models share one amplitude; constraints of the spectrum parameters are applied at the end."""
    from .component import CombinedComponent
    observations = _as_tuple(observations)
    model_frame = observations[0].model_frame
    parameters, slot, models = [], [], []
    for i, src in enumerate(sources):
        for j, comp in enumerate(src.children if isinstance(src, CombinedComponent) else (src,)):
            p = comp.get_parameter("spectrum")
            parameters.append(p)
            if p is not None and not p.fixed:
                p[:] = 1  # flat spectrum: the rendered model is the morphology seen through every channel
            model = comp.get_model(frame=model_frame)
            for known, other in enumerate(models):  # identical initial models make the normal equations singular
                if np.allclose(model, other):
                    logger.warning("Source %d, component %d has a model identical to another component; their spectra will be "
                                   "identical (the duplicate should probably be removed)", i, j)
                    slot.append(known)
                    break
            else:
                slot.append(len(models))
                models.append(model)
    n_models = len(models)
    for obs in observations:
        rendered = np.stack([np.asarray(obs.render(model)) for model in models], axis=0)  # (n_models, C, Ny, Nx)
        amplitudes = np.zeros((n_models, obs.C))
        for c in range(obs.C):
            d = np.asarray(obs.data[c]).reshape(-1)
            w = np.asarray(obs.weights[c]).reshape(-1)
            m = rendered[:, c].reshape(n_models, -1)
            mw = m * w[None, :]
            # convolution spreads flux beyond the box: compare weighted with unweighted flux to see whether most of a
            # component's flux in this channel sits on pixels that carry weight at all
            use = np.flatnonzero(np.sum(mw, axis=1) / np.sum(m, axis=1) / np.mean(w) > 0.1)
            if len(use) == n_models:
                amplitudes[:, c] = np.linalg.inv(mw @ m.T) @ m @ (d * w)
            else:
                amplitudes[use, c] = np.linalg.inv(mw[use] @ m[use].T) @ m[use] @ (d * w)
        for p, k in zip(parameters, slot):
            if p is not None and not p.fixed:
                obs.renderer.map_channels(p)[:] = amplitudes[k]
    for p in parameters:
        if p is not None and p.constraint is not None:
            p[:] = p.constraint(p, 0)
