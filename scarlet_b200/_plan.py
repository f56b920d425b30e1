"""Host "compiler": walks Blend objects and emits the flat batch descriptor the CUDA plan consumes; moves
parameters / optimiser state between the host ``Parameter`` objects and the device.

This is the seam the reference fills with closures handed to ``proxmin.adaprox`` (scarlet/blend.py:103-180).
Anything the device cannot express (unknown Constraint / Renderer / Morphology classes, priors, callable steps
other than ``relative_step``) raises here -- there is no host fallback.
"""
import ctypes
from functools import partial

import numpy as np
import numpy.ma as ma

from . import _native as nat
from .component import CombinedComponent, FactorizedComponent
from .constraint import MonoTables, chain_desc, constraint_ops
from .morphology import ImageMorphology, PointSourceMorphology
from .parameter import _StateLink, relative_step
from .psf import GaussianPSF
from .renderer import ConvolutionRenderer, NullRenderer, ResolutionRenderer
from .spectrum import TabulatedSpectrum


def leaf_components(sources):
    """Depth-first list of FactorizedComponents (the order of ``Blend.parameters``, model.py:51-54)."""
    out = []
    for s in sources:
        if isinstance(s, FactorizedComponent):
            out.append(s)
        elif isinstance(s, CombinedComponent):
            if getattr(s, "operation", "add") != "add":
                raise TypeError("CombinedComponent(operation=%r) is not on the device path" % s.operation)
            out += leaf_components(s.children)
        else:
            raise TypeError("source of type %s is not on the device path (FactorizedComponent / CombinedComponent "
                            "of FactorizedComponents only)" % type(s).__name__)
    return out


import threading  # noqa: E402

_REUSE_STEPS = threading.local()  # .on: inside replace_sources of this thread


def _step_of_spectrum(p, C):
    """-> (factor, min_step[C]); factor < 0 encodes a constant scalar step."""
    st = p.step
    # inside one fit (re-plans of dynamic boxes) the descriptor of an unchanged step object is reused; a new fit reads it afresh
    cached = p.__dict__.get("_sb_step") if (getattr(_REUSE_STEPS, "on", False) and hasattr(p, "__dict__")) else None
    if cached is not None and cached[0] is st and cached[1] == C:
        return cached[2], cached[3].copy()
    out = _step_of_spectrum_uncached(p, st, C)
    if hasattr(p, "__dict__"):
        p.__dict__["_sb_step"] = (st, C, out[0], out[1].copy())
    return out


def _step_of_spectrum_uncached(p, st, C):
    if isinstance(st, partial) and st.func is relative_step and not st.args:
        kw = dict(st.keywords)
        factor = float(kw.pop("factor", 0.1))
        minimum = kw.pop("minimum", 0)
        if kw.pop("axis", None) is not None or kw:
            raise TypeError("relative_step with axis/extra keywords is not on the device path")
        mn = np.array(ma.filled(ma.asarray(minimum, dtype=np.float64), np.inf), dtype=np.float64)
        return factor, np.broadcast_to(mn, (C,)).copy()
    if st is relative_step:
        return 0.1, np.zeros(C)
    if callable(st):
        raise TypeError("callable step %r of parameter '%s' has no device equivalent (only relative_step)" % (st, p.name))
    return -1.0, np.full(C, float(st))


def _const_step(p):
    if callable(p.step):
        raise TypeError("parameter '%s': only constant steps are supported on the device for this parameter kind "
                        "(got %r)" % (p.name, p.step))
    return float(p.step)


class _HostStore:
    """Packed float64 host images (pinned) of every parameter of a plan: value and optimiser state."""
    KEYS = ("value", "m", "v", "vhat")

    def __init__(self):
        self.arrays = {}
        self.valid = False          # m/v/vhat hold the device state of the last fit
        self.device_synced = False  # ... and the device still holds exactly that state


class DevicePlan:
    """One CUDA plan for a batch of structurally identical scenes (same frame / observation shapes)."""

    def __init__(self, blends, precision=32, device=None):
        self.blends = list(blends)
        self.precision = int(precision)
        self.device = nat.default_device() if device is None else int(device)
        self._handle = None
        self._keep = []  # host arrays referenced by the descriptor until sb_plan_create returns
        self._host_obs = None
        self._pinned_ptrs = []
        b0 = self.blends[0]
        self.frame_shape = tuple(b0.frame.shape)
        C, Ny, Nx = self.frame_shape
        if C > nat.SB_MAX_CHANNELS:
            raise ValueError("at most %d channels" % nat.SB_MAX_CHANNELS)
        self.S = len(self.blends)

        desc = nat.sb_batch_desc()
        desc.precision, desc.n_scenes, desc.C, desc.Ny, desc.Nx = self.precision, self.S, C, Ny, Nx

        # ---- observations ------------------------------------------------------------------------
        n_obs = len(b0.observations)
        if not 1 <= n_obs <= nat.SB_MAX_OBS:
            raise ValueError("between 1 and %d observations per scene" % nat.SB_MAX_OBS)
        desc.n_obs = n_obs
        self.obs_meta = []
        for o in range(n_obs):
            metas = [self._obs_meta(b, o) for b in self.blends]
            m0 = metas[0]
            for m in metas[1:]:
                if (m["kind"], m["shape"], m["chan_off"], m["origin"], m["fshape"], m["korigin"], m["kshape"]) != \
                        (m0["kind"], m0["shape"], m0["chan_off"], m0["origin"], m0["fshape"], m0["korigin"], m0["kshape"]):
                    raise ValueError("all scenes of a batch need identically shaped observations")
            shared = all(m["renderer"] is m0["renderer"] for m in metas)
            has_shift = m0["psf_shift"] is not None
            if any((m["psf_shift"] is not None) != has_shift or m["shift_grid"] != m0["shift_grid"] for m in metas):
                raise ValueError("all scenes of a batch need the same renderer structure (psf_shift)")
            if has_shift and shared and self.S > 1:
                raise ValueError("psf_shift is fitted per scene: give every scene its own ConvolutionRenderer")
            desc.obs[o] = nat.sb_obs_desc(m0["kind"], m0["shape"][0], m0["shape"][1], m0["shape"][2], m0["chan_off"],
                                          m0["origin"][0], m0["origin"][1], m0["fshape"][0], m0["fshape"][1], int(shared),
                                          int(has_shift), m0["shift_grid"][0], m0["shift_grid"][1],
                                          int(bool(m0["psf_shift"].fixed)) if has_shift else 0,
                                          _const_step(m0["psf_shift"]) if has_shift else 0.0)
            self.obs_meta.append(dict(metas=metas, shared=shared))

        self._desc = desc
        self._describe_sources(desc)
        handle = ctypes.c_void_p()
        nat.check(nat.lib().sb_plan_create(ctypes.byref(desc), self.device, ctypes.byref(handle)))
        self._handle = handle
        self._build_store()
        self.upload_observations()

    def replace_sources(self):
        """Dynamic boxes: the sources of the same scenes changed shape (``ImageMorphology.update``).  Only the source side of
        the device plan is rebuilt (``sb_plan_set_sources``); data, weights, K^ and every spectral buffer stay on the device.
        Parameters and optimiser state travel through the host: upload them afterwards."""
        self._detach_store()
        for p in getattr(self, "_store_ptrs", []):  # back to the pool (still owned by the plan, freed in close())
            self._pinned_pool.append((p, self._pinned_sizes[p]))
        self._store_ptrs = []
        _REUSE_STEPS.on = True
        try:
            self._describe_sources(self._desc)
        finally:
            _REUSE_STEPS.on = False
        nat.check(nat.lib().sb_plan_set_sources(self._handle, ctypes.byref(self._desc)))
        self._replanning = True
        try:
            self._build_store()
        finally:
            self._replanning = False

    def _describe_sources(self, desc):
        """Walk the sources of every scene -> flat source / chain / table descriptors inside ``desc`` (kept alive in ``_keep``)."""
        C, Ny, Nx = self.frame_shape
        tables = MonoTables()
        chains, chain_keys = [], {}

        def chain_index(constraint, shape):
            if constraint is None:
                return -1
            ops, repeat = constraint_ops(constraint, shape, tables)
            key = (tuple(ops), repeat)
            if key not in chain_keys:
                chain_keys[key] = len(chains)
                chains.append(chain_desc(ops, repeat))
            return chain_keys[key]

        self.slots = []  # per leaf component: dict(kind, spectrum, image|center, shift)
        starts = [0]
        psf_box, psf_sigma = 0, np.zeros(nat.SB_MAX_CHANNELS)
        src_descs = []
        for b in self.blends:
            if tuple(b.frame.shape) != self.frame_shape:
                raise ValueError("all scenes of a batch need the same model frame shape")
            for comp in leaf_components(b.sources):
                spec, morph = comp.children[0], comp.children[1]
                if not isinstance(spec, TabulatedSpectrum):
                    raise TypeError("spectrum model %s is not on the device path" % type(spec).__name__)
                sp = spec.parameters[0]
                if sp.shape != (C,) or spec.bbox.origin != (0,):
                    raise NotImplementedError("spectra must cover all model channels")
                for p in comp.parameters:
                    if p.prior is not None:
                        raise NotImplementedError("priors are not supported on the device path (parameter '%s')" % p.name)
                d = nat.sb_source_desc()
                factor, mn = _step_of_spectrum(sp, C)
                d.sed_step_factor = factor
                for c in range(C):
                    d.sed_step_min[c] = mn[c]
                d.sed_chain = chain_index(sp.constraint, (1, C))
                d.sed_is_f32 = int(sp.dtype == np.float32)
                d.sed_fixed = int(bool(sp.fixed))
                slot = dict(spectrum=sp, comp=comp)
                if isinstance(morph, PointSourceMorphology):
                    psf = morph.psf
                    if not isinstance(psf, GaussianPSF) or not psf.integrate:
                        raise TypeError("PointSourceMorphology needs a pixel-integrated GaussianPSF model PSF on the device path")
                    sig = np.asarray(psf.sigma, dtype=np.float64)
                    sig = np.full(C, sig[0]) if sig.size == 1 else sig
                    if sig.size != C:
                        raise ValueError("model PSF sigma must have one entry per channel")
                    bs = psf.bbox.shape[1]
                    if psf_box and (psf_box != bs or not np.array_equal(psf_sigma[:C], sig)):
                        raise ValueError("all point sources of a batch must share the model PSF")
                    psf_box, psf_sigma[:C] = bs, sig
                    cen = morph.parameters[0]
                    d.kind, d.By, d.Bx = 1, bs, bs
                    d.oy, d.ox = morph.bbox.origin[-2], morph.bbox.origin[-1]
                    d.chain = -1
                    if cen.constraint is not None:
                        raise TypeError("constraints on point-source centres are not on the device path")
                    d.morph_step = _const_step(cen)
                    d.morph_fixed = int(bool(cen.fixed))
                    slot.update(kind=1, center=cen)
                elif isinstance(morph, ImageMorphology):
                    img = morph.parameters[0]
                    d.kind, d.By, d.Bx = 0, img.shape[0], img.shape[1]
                    d.oy, d.ox = morph.bbox.origin[-2], morph.bbox.origin[-1]
                    d.chain = chain_index(img.constraint, img.shape)
                    d.morph_step = _const_step(img)
                    d.morph_fixed = int(bool(img.fixed))
                    d.resizing = int(bool(getattr(morph, "resizing", False)) and not img.fixed)
                    shift = morph.parameters[1] if len(morph.parameters) > 1 else None
                    if morph.shifting:  # the shift is a fitted parameter: it travels in the centre arrays
                        from . import fft
                        if shift is None or shift.constraint is not None:
                            raise TypeError("a shifting morphology needs an unconstrained 'shift' parameter")
                        fshape = fft._get_fft_shape(img._data, img._data, padding=10, axes=(0, 1))
                        d.shifting, d.shift_Fy, d.shift_Fx = 1, int(fshape[0]), int(fshape[1])
                        d.shift_step = 0.0 if shift.fixed else _const_step(shift)
                        slot.update(kind=0, image=img, shift=None, center=shift, morph=morph)
                    else:
                        slot.update(kind=0, image=img, shift=shift, morph=morph)
                else:
                    raise TypeError("morphology model %s is not on the device path" % type(morph).__name__)
                slot["scene"] = len(starts) - 1
                src_descs.append(d)
                self.slots.append(slot)
            starts.append(len(src_descs))
        self.n_src = len(src_descs)
        self.C = C
        self.ext = [s for s in self.slots if s["kind"] == 0]
        self.pts = [s for s in self.slots if s.get("center") is not None]  # point-source centres and image shifts
        for om in self.obs_meta:  # renderer parameters (psf_shift): behind the sources', observation-major, one per scene
            for m in om["metas"]:
                if m["psf_shift"] is not None:
                    self.pts.append(dict(center=m["psf_shift"]))
        self.morph_sizes = [s["image"].size for s in self.ext]
        self.morph_offsets = np.concatenate([[0], np.cumsum(self.morph_sizes)]).astype(np.int64)
        self.n_morph = int(self.morph_offsets[-1])

        src_arr = (nat.sb_source_desc * max(self.n_src, 1))(*src_descs)
        chain_arr = (nat.sb_chain_desc * max(len(chains), 1))(*chains)
        mono_arr = tables.descs()
        start_arr = np.asarray(starts, dtype=np.int32)
        desc.n_sources, desc.n_chains, desc.n_mono = self.n_src, len(chains), len(tables.tables)
        desc.psf_boxsize = psf_box
        for c in range(nat.SB_MAX_CHANNELS):
            desc.psf_sigma[c] = psf_sigma[c]
        desc.scene_src_start = nat.ptr(start_arr)
        desc.sources = ctypes.addressof(src_arr)
        desc.chains = ctypes.addressof(chain_arr)
        desc.mono = ctypes.addressof(mono_arr)
        self._keep = [src_arr, chain_arr, mono_arr, start_arr, tables]

    def inspect(self):
        """Dynamic boxes: the device's reading of ``ImageMorphology.update``'s rules for every source of a paused scene ->
        int32 array over the sources: new box size, 0 = keep, -1 = the host has to decide."""
        action = np.zeros(max(self.n_src, 1), dtype=np.int32)
        nat.check(nat.lib().sb_plan_inspect(self._handle, nat.ptr(action)))
        return action[:self.n_src]

    # -------------------------------------------------------------------------------------------------
    @staticmethod
    def _obs_meta(blend, o):
        obs = blend.observations[o]
        r = getattr(obs, "renderer", None)
        if r is None:
            raise RuntimeError("observation %d is not matched to the model frame (call obs.match(frame))" % o)
        korigin, kernel, operator = (0, 0), None, None
        if type(r) is ConvolutionRenderer:
            fshape, korigin, kernel = r.device_kernel()
            kind = 0
        elif type(r) is NullRenderer:
            fshape, kind = (blend.frame.shape[1], blend.frame.shape[2]), 1
        elif type(r) is ResolutionRenderer:
            operator = r.device_operator()
            fshape, kind = operator["fshape"], 3 if operator["rotated"] else 2
            if kind == 3 and max(obs.data.shape[1:]) > 32:
                raise NotImplementedError("a rotated ResolutionRenderer is on the device path for low-resolution cubes up to 32x32 "
                                          "pixels (got %s)" % (tuple(obs.data.shape[1:]),))
            if nat.lib().sb_fft_supported_length(int(fshape[0])) != fshape[0] or nat.lib().sb_fft_supported_length(int(fshape[1])) != fshape[1]:
                raise NotImplementedError("ResolutionRenderer grid %s: the Fourier interpolation is tied to the reference's grid, and the "
                                          "fused spectral kernels are not instantiated for these lengths" % (tuple(fshape),))
        else:
            raise TypeError("renderer %s is not on the device path (ConvolutionRenderer / NullRenderer / ResolutionRenderer)"
                            % type(r).__name__)
        psf_shift = None
        if r.parameters:
            if kind != 0 or len(r.parameters) != 1 or r.parameters[0].name != "psf_shift":
                raise NotImplementedError("renderer parameters other than ConvolutionRenderer's psf_shift are not on the device path")
            psf_shift = r.parameters[0]
            if psf_shift.constraint is not None or psf_shift.prior is not None or psf_shift.shape != (2,):
                raise TypeError("psf_shift must be an unconstrained (dy, dx) parameter")
        return dict(kind=kind, shape=tuple(obs.data.shape), chan_off=r.channel_offset, origin=tuple(r.origin),
                    fshape=tuple(int(f) for f in fshape), kernel=kernel, korigin=tuple(korigin),
                    kshape=None if kernel is None else kernel.shape, obs=obs, renderer=r, operator=operator, psf_shift=psf_shift,
                    shift_grid=r.shift_grid() if psf_shift is not None else (0, 0))

    def _pinned(self, shape, dtype):
        """numpy view of pinned host memory (sb_host_alloc) -- staging for asynchronous H2D copies.  Buffers returned to the
        pool by ``replace_sources`` are reused when large enough (cudaHostAlloc costs milliseconds)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        pool = self.__dict__.setdefault("_pinned_pool", [])
        fit = [i for i, (q, size) in enumerate(pool) if size >= max(n, 8)]
        if fit:
            i = min(fit, key=lambda j: pool[j][1])
            p, size = pool.pop(i)
            self.__dict__.setdefault("_pinned_sizes", {})[p] = size
        else:
            size = max(n, 8)
            if self.__dict__.get("_replanning"):
                size = int(size * 1.25)  # a re-planned batch: boxes tend to grow, leave headroom for the next re-plan
            p = nat.lib().sb_host_alloc(size)
            if not p:
                raise MemoryError("sb_host_alloc(%d) failed" % size)
            self._pinned_ptrs.append(p)
            self.__dict__.setdefault("_pinned_sizes", {})[p] = size
        buf = (ctypes.c_byte * max(n, 8)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def _stage_observations(self, reuse=None):
        """Stack the per-scene cubes once into pinned staging buffers (host work, not repeated per upload)."""
        C, Ny, Nx = self.frame_shape
        self._host_obs = []
        for o, om in enumerate(self.obs_meta):
            if reuse is not None:  # same pinned buffers, new contents of data / weights / constants
                h = reuse[o]
                for i, m in enumerate(om["metas"]):
                    h["data"][i] = m["obs"].data
                    h["weights"][i] = m["obs"].weights
                h["consts"][...] = self._loss_consts(om["metas"])
                self._host_obs.append(h)
                continue
            metas = om["metas"]
            shp = (self.S,) + metas[0]["shape"]
            # cubes travel in the plan's precision: a float64 twin must not see data rounded to float32
            dt = np.float64 if self.precision == 64 else np.float32
            data, weights = self._pinned(shp, dt), self._pinned(shp, dt)
            for i, m in enumerate(metas):
                data[i] = m["obs"].data
                weights[i] = m["obs"].weights
            kernels = None
            if metas[0]["kind"] == 0:
                ks = [metas[0]["kernel"]] if om["shared"] else [m["kernel"] for m in metas]
                kernels = self._pinned((len(ks),) + ks[0].shape, np.float64)
                for i, k in enumerate(ks):
                    kernels[i] = k
            resamp = None
            if metas[0]["kind"] in (2, 3):
                ops = [metas[0]["operator"]] if om["shared"] else [m["operator"] for m in metas]
                tables = ("A", "B") if metas[0]["kind"] == 3 else ("Ey", "Ex")
                for op in ops[1:]:
                    if not (all(np.array_equal(op[t], ops[0][t]) for t in tables) and op["scale"] == ops[0]["scale"]):
                        raise ValueError("all scenes of a batch need the same resampling geometry")
                khat = self._pinned((len(ops),) + ops[0]["khat"].shape, np.complex128)
                for i, op in enumerate(ops):
                    khat[i] = op["khat"]
                Fy, Fx = ops[0]["fshape"]
                resamp = dict(khat=khat, tables=tuple(ops[0][t] for t in tables), rotated=metas[0]["kind"] == 3,
                              h2=ops[0]["scale"] * Fy * Fx)
            self._host_obs.append(dict(data=data, weights=weights, kernels=kernels, korigin=metas[0]["korigin"], resamp=resamp,
                                       consts=self._loss_consts(metas)))

    def _loss_consts(self, metas):
        """per scene: log_norm (observation.py:172-186) + the chi^2 of data pixels the model frame does not cover"""
        C, Ny, Nx = self.frame_shape
        consts = []
        for m in metas:
            obs, (oy, ox) = m["obs"], m["origin"]
            if m["kind"] in (2, 3):  # resampled observation: every pixel is rendered
                consts.append(float(obs.log_norm))
                continue
            H, W = obs.data.shape[1:]
            outside = np.ones((H, W), dtype=bool)
            y0, y1, x0, x1 = max(0, -oy), min(H, Ny - oy), max(0, -ox), min(W, Nx - ox)
            if y1 > y0 and x1 > x0:
                outside[y0:y1, x0:x1] = False
            extra = 0.0
            if outside.any():
                w = np.asarray(obs.weights, dtype=np.float64)[:, outside]
                dd = np.asarray(obs.data, dtype=np.float64)[:, outside]
                extra = 0.5 * float((w * dd * dd).sum())
            consts.append(float(obs.log_norm) + extra)
        return np.asarray(consts, dtype=np.float64)

    def upload_observations(self):
        """Host (pinned) -> device copy of data, weights, difference-kernel images and the per-scene loss constants;
        the device transforms the kernels to K^ itself (as fft.convolve does on every call, fft.py:385-388)."""
        if self._host_obs is None:
            self._stage_observations()
        nbytes = 0
        for o, h in enumerate(self._host_obs):
            ker, rs = h["kernels"], h["resamp"]
            up = nat.lib().sb_plan_upload_observation_f64 if h["data"].dtype == np.float64 else nat.lib().sb_plan_upload_observation
            nat.check(up(self._handle, o, nat.ptr(h["data"]), nat.ptr(h["weights"]),
                         nat.ptr(rs["khat"].view(np.float64)) if rs is not None else None, nat.ptr(h["consts"])))
            if rs is not None:
                up_rs = nat.lib().sb_plan_upload_resampling_rot if rs["rotated"] else nat.lib().sb_plan_upload_resampling
                t0, t1 = rs["tables"]
                nat.check(up_rs(self._handle, o, nat.ptr(t0.view(np.float64)), nat.ptr(t1.view(np.float64)), rs["h2"]))
                nbytes += rs["khat"].nbytes + t0.nbytes + t1.nbytes
            if ker is not None:
                nat.check(nat.lib().sb_plan_upload_kernels(self._handle, o, nat.ptr(ker), ker.shape[-2], ker.shape[-1],
                                                           h["korigin"][0], h["korigin"][1]))
            nbytes += h["data"].nbytes + h["weights"].nbytes + (ker.nbytes if ker is not None else 0) + h["consts"].nbytes
        return nbytes

    def refresh_observations(self):
        """Re-read ``obs.data`` / ``obs.weights`` of every scene (they may have been edited in place since the plan was built, and
        the reference reads them afresh on every fit) and upload them with the loss constants; kernels are not touched.
        ``log_norm`` / ``noise_rms`` stay cached on the Observation exactly as in the reference (observation.py:116-124, 172-186)."""
        if self._host_obs is None:
            return self.upload_observations()
        old, self._host_obs = self._host_obs, None
        self._stage_observations(reuse=old)
        nbytes = 0
        for o, h in enumerate(self._host_obs):
            up = nat.lib().sb_plan_upload_observation_f64 if h["data"].dtype == np.float64 else nat.lib().sb_plan_upload_observation
            nat.check(up(self._handle, o, nat.ptr(h["data"]), nat.ptr(h["weights"]), None, nat.ptr(h["consts"])))
            nbytes += h["data"].nbytes + h["weights"].nbytes + h["consts"].nbytes
        return nbytes

    # -------------------------------------------------------------------------------------------------
    def _build_store(self):
        """Pinned packed host images of the parameters (value, m, v, vhat) + pointer tables into the Parameters."""
        nsed, nmorph, ncen = max(self.n_src, 1), max(self.n_morph, 1), max(len(self.pts), 1)
        self.store = _HostStore()
        for key in _HostStore.KEYS:
            self.store.arrays[key] = dict(sed=self._pinned((nsed, self.C), np.float64), morph=self._pinned((nmorph,), np.float64),
                                          center=self._pinned((ncen, 2), np.float64))
            for a in self.store.arrays[key].values():
                a[...] = 0
        # (parameter, link) in device order, per group
        self._linked = []
        for k, s in enumerate(self.slots):
            self._linked.append((s["spectrum"], _StateLink(self.store, "sed", k * self.C, (k + 1) * self.C, s["spectrum"].shape)))
        for s, a, b in zip(self.ext, self.morph_offsets[:-1], self.morph_offsets[1:]):
            self._linked.append((s["image"], _StateLink(self.store, "morph", int(a), int(b), s["image"].shape)))
        for i, s in enumerate(self.pts):
            self._linked.append((s["center"], _StateLink(self.store, "center", 2 * i, 2 * i + 2, s["center"].shape)))
        self._unused = [s["shift"] for s in self.ext if s.get("shift") is not None]
        # pointer tables for the C gather/scatter of the values (float32/float64 contiguous parameters)
        self._tables = {}
        for group in ("sed", "morph", "center"):
            params = [(p, l) for p, l in self._linked if l.group == group]
            ok = all(p.flags.c_contiguous and p.dtype in (np.float32, np.float64) and p.size == l.stop - l.start for p, l in params)
            if ok and params:
                ptrs = np.array([p.ctypes.data for p, _ in params], dtype=np.uint64)
                counts = np.array([p.size for p, _ in params], dtype=np.int64)
                f32 = np.array([p.dtype == np.float32 for p, _ in params], dtype=np.int32)
                self._tables[group] = (ptrs, counts, f32, len(params))
            else:
                self._tables[group] = None
        self._param_refs = [p for p, _ in self._linked]  # keep the arrays (and their addresses) alive
        self._store_ptrs = [a.ctypes.data for key in _HostStore.KEYS for a in self.store.arrays[key].values()]

    def _gather_values(self):
        vals = self.store.arrays["value"]
        for group in ("sed", "morph", "center"):
            t = self._tables[group]
            dst = vals[group].reshape(-1)
            if t is not None:
                nat.check(nat.lib().sb_host_gather_f64(nat.ptr(dst), nat.ptr(t[0]), nat.ptr(t[1]), nat.ptr(t[2]), t[3]))
            else:
                for p, l in self._linked:
                    if l.group == group:
                        dst[l.start:l.stop] = np.asarray(p._data, dtype=np.float64).reshape(-1)

    def _scatter_values(self):
        vals = self.store.arrays["value"]
        for group in ("sed", "morph", "center"):
            t = self._tables[group]
            src = vals[group].reshape(-1)
            if t is not None:
                nat.check(nat.lib().sb_host_scatter_f64(nat.ptr(src), nat.ptr(t[0]), nat.ptr(t[1]), nat.ptr(t[2]), t[3]))
            else:
                for p, l in self._linked:
                    if l.group == group:
                        p._data[...] = src[l.start:l.stop].reshape(p.shape)

    def _upload(self, key):
        a = self.store.arrays[key]
        nat.check(nat.lib().sb_plan_upload_params(self._handle, _HostStore.KEYS.index(key), nat.ptr(a["sed"]), nat.ptr(a["morph"]),
                                                  nat.ptr(a["center"])))

    def upload_parameters(self, state=True):
        """Host Parameters -> device.  Values are packed by one C call; optimiser state that still lives in this
        plan's packed arrays (the result of the previous fit) is uploaded as is; a cold start (no m/v/vhat anywhere)
        zeroes the state on the device instead of shipping zeros (blend.py:154-163)."""
        self._gather_values()
        self._upload("value")
        nbytes = sum(a.nbytes for a in self.store.arrays["value"].values())
        if not state:
            return nbytes
        store = self.store
        cold = not store.valid
        touched = False
        for p, link in self._linked:
            d = p.__dict__
            own = d.get("_link")
            if own is not None and own.store is not store and own.store.valid:  # state lives in another plan: adopt it
                for key in ("m", "v", "vhat"):
                    if d.get("_" + key) is None:
                        d["_" + key] = own.view(key)
            for key in ("m", "v", "vhat"):
                val = d.get("_" + key)
                if val is not None:
                    if cold:  # first explicit state: the packed arrays become the truth (zeros elsewhere)
                        for k2 in ("m", "v", "vhat"):
                            for a in store.arrays[k2].values():
                                a[...] = 0
                        cold = False
                        store.valid = True
                    store.arrays[key][link.group].reshape(-1)[link.start:link.stop] = np.asarray(val, dtype=np.float64).reshape(-1)
                    d["_" + key] = None
                    touched = True
            d["_link"] = link
        if cold:
            nat.check(nat.lib().sb_plan_zero_state(self._handle))
            store.device_synced = False
            return nbytes
        if store.device_synced and not touched:
            return nbytes  # the device still holds exactly this state (left there by the previous fit)
        for key in ("m", "v", "vhat"):
            self._upload(key)
        store.device_synced = True
        return 4 * nbytes

    def upload_values(self, params, values):
        """Upload explicit values for the given Parameter objects (same order) instead of their stored contents; the host
        Parameters are not touched.  Backs ``Blend.get_model(*parameters)``."""
        given = {id(p): np.asarray(v, dtype=np.float64) for p, v in zip(params, values)}
        vals = self.store.arrays["value"]
        for p, l in self._linked:
            v = given.get(id(p))
            if v is None:
                v = np.asarray(p._data, dtype=np.float64)
            if v.size != l.stop - l.start:
                raise ValueError("parameter '%s': expected %d values, got shape %s" % (p.name, l.stop - l.start, v.shape))
            vals[l.group].reshape(-1)[l.start:l.stop] = v.reshape(-1)
        self._upload("value")

    def forget_state(self, values=None):
        """Benchmark helper: drop the optimiser state everywhere (next fit is a cold start) and optionally restore the
        parameter values from a ``pack_current()`` snapshot."""
        self.store.valid = False
        self.store.device_synced = False
        for p, link in self._linked:
            d = p.__dict__
            d["_m"] = d["_v"] = d["_vhat"] = d["_std"] = None
        if values is not None:
            for g, a in zip(("sed", "morph", "center"), values):
                self.store.arrays["value"][g][...] = a
            self._scatter_values()

    def download_parameters(self, state=True):
        """Device -> the same host Parameter objects: values in place (dtype preserved); m/v/vhat as views into the
        plan's packed float64 arrays and std = 1/sqrt(masked v) (blend.py:189-192), both materialised when read."""
        keys = _HostStore.KEYS if state else ("value",)
        for which, key in enumerate(keys):
            a = self.store.arrays[key]
            nat.check(nat.lib().sb_plan_download_params(self._handle, which, nat.ptr(a["sed"]), nat.ptr(a["morph"]), nat.ptr(a["center"])))
        self._scatter_values()
        nbytes = sum(a.nbytes for a in self.store.arrays["value"].values()) * len(keys)
        if state:
            self.store.valid = True
            self.store.device_synced = True
            for p, link in self._linked:
                d = p.__dict__
                d["_m"] = d["_v"] = d["_vhat"] = d["_std"] = None
                d["_link"] = link
            for sh in self._unused:  # free but unused parameter (morphology.py:112-113): zero gradient forever
                d = sh.__dict__
                if d.get("_v") is None:
                    d["_m"], d["_v"], d["_vhat"] = np.zeros(sh.shape), np.zeros(sh.shape), np.zeros(sh.shape)
                d["_std"], d["_std_from_v"] = None, True  # std = 1/sqrt(masked v), evaluated when read
        return nbytes

    def pack_current(self):
        """Snapshot (copies) of the packed host arrays: value, m, v, vhat -> tuples (sed, morph, center)."""
        self._gather_values()
        return [tuple(self.store.arrays[key][g].copy() for g in ("sed", "morph", "center")) for key in _HostStore.KEYS]

    # -------------------------------------------------------------------------------------------------
    def evaluate(self, obs=0, want=("model", "rendered", "loss", "grads")):
        """One forward + backward at the device's current parameters (no update)."""
        C, Ny, Nx = self.frame_shape
        m0 = self.obs_meta[obs]["metas"][0]
        out = {}
        model = np.zeros((self.S, C, Ny, Nx)) if "model" in want else None
        rendered = np.zeros((self.S,) + m0["shape"]) if "rendered" in want else None
        loss = np.zeros(self.S) if "loss" in want else None
        g_sed = g_morph = g_cen = None
        if "grads" in want:
            g_sed = np.zeros((max(self.n_src, 1), self.C))
            g_morph = np.zeros(max(self.n_morph, 1))
            g_cen = np.zeros((max(len(self.pts), 1), 2))
        nat.check(nat.lib().sb_plan_evaluate(self._handle, obs, nat.ptr(model), nat.ptr(rendered), nat.ptr(loss),
                                             nat.ptr(g_sed), nat.ptr(g_morph), nat.ptr(g_cen)))
        out.update(model=model, rendered=rendered, loss=loss)
        if "grads" in want:
            out["g_sed"] = g_sed[:self.n_src]
            out["g_morph"] = [g_morph[a:b].reshape(s["image"].shape)
                              for s, a, b in zip(self.ext, self.morph_offsets[:-1], self.morph_offsets[1:])]
            out["g_center"] = g_cen[:len(self.pts)]
        return out

    def fit(self, opts):
        n_iter = np.zeros(self.S, dtype=np.int32)
        status = np.zeros(self.S, dtype=np.int32)
        loss = np.zeros((self.S, max(opts.max_iter, 1)))
        nat.check(nat.lib().sb_plan_fit(self._handle, ctypes.byref(opts), nat.ptr(n_iter), nat.ptr(loss), nat.ptr(status)))
        return n_iter, loss, status

    def fit_enqueue(self, opts, n_iterations):
        nat.check(nat.lib().sb_plan_fit_enqueue(self._handle, ctypes.byref(opts), int(n_iterations)))

    def sync(self):
        nat.check(nat.lib().sb_plan_sync(self._handle))

    def timer_start(self):
        nat.check(nat.lib().sb_plan_timer_start(self._handle))

    def timer_stop(self):
        ms = ctypes.c_float()
        nat.check(nat.lib().sb_plan_timer_stop(self._handle, ctypes.byref(ms)))
        return float(ms.value)

    def profile(self, opts, n_iterations):
        ms = np.zeros(nat.SB_N_STAGES, dtype=np.float32)
        nat.check(nat.lib().sb_plan_profile_iterations(self._handle, ctypes.byref(opts), int(n_iterations), nat.ptr(ms)))
        names = [nat.lib().sb_stage_name(i).decode() for i in range(nat.SB_N_STAGES)]
        return dict(zip(names, (ms / max(n_iterations, 1)).tolist()))

    def prox_histogram(self, enable=True):
        """Diagnostic: counts of proximal sub-iterations executed per source update since the last call (index = count)."""
        out = np.zeros(16, dtype=np.int64)
        nat.check(nat.lib().sb_plan_prox_histogram(self._handle, int(bool(enable)), nat.ptr(out)))
        return out

    @property
    def spectral_mode(self):
        """1: fused row/column spectral kernels, 0: cuFFT pipeline."""
        return int(nat.lib().sb_plan_spectral_mode(self._handle))

    @property
    def kernel_launches(self):
        return int(nat.lib().sb_plan_kernel_launches(self._handle))

    @property
    def device_bytes(self):
        return int(nat.lib().sb_plan_device_bytes(self._handle))

    def device_params(self):
        sed, morph = ctypes.c_void_p(), ctypes.c_void_p()
        n_sed, n_morph, eb = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        nat.check(nat.lib().sb_plan_device_params(self._handle, ctypes.byref(sed), ctypes.byref(n_sed), ctypes.byref(morph),
                                                  ctypes.byref(n_morph), ctypes.byref(eb)))
        return dict(sed=sed.value, n_sed=n_sed.value, morph=morph.value, n_morph=n_morph.value, elem_bytes=eb.value)

    def close(self):
        if self._handle is not None:
            nat.lib().sb_plan_destroy(self._handle)
            self._handle = None
        self._host_obs = None
        self._detach_store()
        for p in self._pinned_ptrs:
            nat.lib().sb_host_free(p)
        self._pinned_ptrs = []

    def _detach_store(self):
        """The pinned arrays are about to be freed: give every linked Parameter private copies of its state."""
        store = getattr(self, "store", None)
        if store is None:
            return
        for p, link in getattr(self, "_linked", []):
            d = p.__dict__
            if d.get("_link") is link:
                if store.valid:
                    for key in ("m", "v", "vhat"):
                        if d.get("_" + key) is None:
                            d["_" + key] = link.view(key).copy()
                    if d.get("_std") is None:
                        d["_std_from_v"] = True
                d["_link"] = None
        store.valid = False
        self.store = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
