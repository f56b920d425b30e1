"""The blended scene and its fitting loop.  Mirrors scarlet/blend.py (``Blend`` 49-308).

``Blend.fit`` keeps the reference's signature, return value and side effects (parameters updated in place,
``m/v/vhat/std`` filled, ``blend.loss`` extended by one entry per gradient evaluation) but the loop itself --
render, PSF convolution, residual, gradients, AMSGrad step, constraint projections, stop rule -- runs on the
GPU as a CUDA graph per iteration (csrc/).  ``BlendBatch`` fits many independent scenes in one plan: the
data-parallel axis the reference leaves to a user-level loop (testing/api.py:216-226).
"""
import logging

import numpy as np

from . import _native as nat
from ._plan import DevicePlan, leaf_components as _leaves
from .component import CombinedComponent
from .model import UpdateException

logger = logging.getLogger("scarlet_b200.blend")


def _fit_options(max_iter, e_rel, min_iter, noise_factor, alg_kwargs, check_every):
    if noise_factor:
        raise NotImplementedError("noise_factor > 0 (host RNG noise injection) is outside the device path")
    kw = dict(alg_kwargs)
    scheme = kw.pop("scheme", "amsgrad")
    if scheme != "amsgrad":
        raise NotImplementedError("only scheme='amsgrad' (the reference default, blend.py:144) is implemented")
    kw.pop("callback", None)  # honoured by Blend.fit (host call per iteration); see there
    prox_max_iter = kw.pop("prox_max_iter", 10)
    b1, b2, eps = kw.pop("b1", 0.9), kw.pop("b2", 0.999), kw.pop("eps", 1e-8)
    kw.pop("p", None)
    if np.ndim(b1) != 0:
        raise NotImplementedError("per-iteration b1 schedules are not supported")
    fixed = bool(kw.pop("fixed_iterations", False))
    if kw:
        raise TypeError("unsupported optimiser keywords: %s" % sorted(kw))
    return nat.fit_opts(max_iter=max_iter, e_rel=e_rel, min_iter=min_iter, prox_max_iter=prox_max_iter,
                        check_every=check_every, fixed_iterations=fixed, b1=b1, b2=b2, eps=eps)


class Blend(CombinedComponent):
    def __init__(self, sources, observations, precision=32, device=None):
        self.sources = sources if hasattr(sources, "__iter__") else (sources,)
        self.observations = observations if hasattr(observations, "__iter__") else (observations,)
        super().__init__(list(self.sources))
        self.loss = []
        self._precision = precision
        self._device = device
        self._plan = None
        self._plan_key = None

    # -- device plan ----------------------------------------------------------------------------------
    def _structure_key(self):
        """Everything the device descriptor bakes in (the reference re-reads all of it on every fit, blend.py:103-145):
        identity, shape, ``fixed``, ``step`` and constraint of every parameter; renderer and cube identity of every
        observation.  In-place edits of ``obs.data`` / ``obs.weights`` are covered by re-staging them in ``fit``."""
        def step_key(p):
            st = p.step
            if callable(st):
                return (id(getattr(st, "func", st)), repr(sorted(getattr(st, "keywords", {}).items(), key=lambda kv: kv[0])))
            return float(st)
        pk = tuple((id(p), p.shape, bool(p.fixed), step_key(p), id(p.constraint), id(p.prior)) for p in self.parameters)
        ok = tuple((id(getattr(o, "renderer", None)), id(o.data), id(o.weights)) + tuple((id(p), bool(p.fixed), float(p.step)) for p in o.parameters)
                   for o in self.observations)
        return pk + ok

    def _get_plan(self, refresh=False):
        key = self._structure_key()
        if self._plan is None or key != self._plan_key:
            if self._plan is not None:
                self._plan.close()
            self._plan = DevicePlan([self], precision=self._precision, device=self._device)
            self._plan_key = key
        elif refresh:
            self._plan.refresh_observations()
        return self._plan

    # -- the fitting loop -----------------------------------------------------------------------------
    def fit(self, max_iter=200, e_rel=1e-3, min_iter=1, noise_factor=0, **alg_kwargs):
        """Fit the model of every source to the data.  Returns ``(len(self.log_likelihood), logL)``.

        Control flow of the reference (blend.py:85-198, 276-302): one ``adaprox`` call runs until ``max_iter`` or
        convergence; after its iterations 10, 20, ... every source may adapt its box (``src.update()``), which aborts the
        call (``UpdateException``) and restarts the optimiser with the new shapes, warm state and ``it = len(self.loss)``.
        Here one ``adaprox`` call is one device plan driven in slices that end exactly at those inspection points."""
        check_every = int(alg_kwargs.pop("check_every", 10))
        # The reference pops scheme / prox_max_iter / callback out of alg_kwargs INSIDE its restart loop (blend.py:143-152):
        # after the first UpdateException they are gone, i.e. a restarted call runs with the defaults and without the user
        # callback.  Mirrored here, not fixed.
        user_cb = alg_kwargs.get("callback", None)
        opts = _fit_options(max_iter, e_rel, min_iter, noise_factor, alg_kwargs, check_every)
        for src in self.sources:
            src.check_parameters()
        if max_iter <= 0:
            return len(self.loss), (-self.loss[-1] if self.loss else None)
        resizing = any(getattr(c.children[1], "resizing", False) and not c.children[1].parameters[0].fixed
                       for c in _leaves(self.sources) if hasattr(c.children[1], "resizing"))
        it = 0
        first = True
        while it < max_iter:
            plan = self._get_plan(refresh=first)
            first = False
            plan.upload_parameters(state=True)
            budget = max_iter - it  # iterations this adaprox call may run
            opts.max_iter, opts.resume, k_done, restarted = budget, 0, 0, False
            n0 = len(self.loss)
            X = self.parameters + tuple(p for obs in self.observations for p in obs.parameters)
            while k_done < budget:
                # without resizable boxes nobody inspects the sources: run the whole call in one go; a user callback sees
                # the parameters after every iteration (blend.py:301-302), so the call is then driven one iteration at a time
                if user_cb is not None:
                    stop = k_done + 1
                else:
                    stop = budget if not resizing else min(budget, 11 if k_done == 0 else k_done + 10)
                opts.run_until = stop
                n_iter, loss, status = plan.fit(opts)
                opts.resume = 1
                n = int(n_iter[0])  # gradient evaluations of this call so far
                plan.download_parameters(state=True)
                self.loss[n0:] = loss[0, :n].tolist()
                if status[0] == nat.SB_ERR_NONFINITE:
                    for src in self.sources:
                        src.check_parameters()  # raises ArithmeticError naming the parameter (model.py:153-165)
                    raise ArithmeticError("a parameter became non-finite during the fit")
                k_done = stop
                last = n - 1  # index of the last iteration of this call (proxmin's ``it``)
                if resizing and n == stop and last > 0 and last % 10 == 0:
                    changed = False
                    for src in self.sources:
                        try:
                            src.update()
                        except UpdateException:
                            changed = True
                    if changed:  # blend.py:196-198
                        it = len(self.loss)
                        restarted = True
                        break
                if n < stop:  # converged inside the slice (StopIteration in the reference)
                    break
                if user_cb is not None:
                    # order of Blend._callback (blend.py:276-302): the stop rule comes before the user callback
                    if (not opts.fixed_iterations and last > min_iter and len(self.loss) >= 2
                            and abs(self.loss[-1] - self.loss[-2]) < e_rel * abs(self.loss[-1])):
                        break
                    before = [np.array(x._data if hasattr(x, "_data") else x, copy=True) for x in X]
                    try:
                        user_cb(*X, it=last)
                    except StopIteration:  # "clean return from proxmin"
                        break
                    if any(not np.array_equal(b, np.asarray(x)) for b, x in zip(before, X)):
                        plan.upload_parameters(state=False)  # the callback edited parameters in place
            if not restarted:
                break
            user_cb = None
            opts.prox_max_iter = 10
        logger.info("scarlet ran for {0} iterations to logL = {1}".format(len(self.loss), -self.loss[-1]))
        return len(self.loss), -self.loss[-1]

    # -- forward --------------------------------------------------------------------------------------
    def get_model(self, *parameters, frame=None):
        """Model of the entire blend in the model frame, rendered on the device."""
        plan = self._get_plan()
        if parameters:  # values to use instead of the stored ones, in the order of ``self.parameters`` (blend.py:200-244)
            if len(parameters) != len(self.parameters):
                raise ValueError("expected %d parameter arrays, got %d" % (len(self.parameters), len(parameters)))
            plan.upload_values(self.parameters, parameters)
        else:
            plan.upload_parameters(state=False)
        model = plan.evaluate(want=("model",))["model"][0].astype(self.frame.dtype)
        if frame is not None and frame is not self.frame and frame.bbox != self.frame.bbox:
            from .bbox import overlapped_slices
            out = np.zeros(frame.shape, dtype=frame.dtype)
            fs, ms = overlapped_slices(frame.bbox, self.frame.bbox)
            out[fs] = model[ms]
            return out
        return model

    @property
    def log_likelihood(self):
        return -np.array(self.loss)

    @property
    def bbox(self):
        return self.frame.bbox


class BlendBatch:
    """Many independent, structurally identical scenes fitted together on one GPU.

    ``n_streams > 1`` splits the batch into that many plans (each with its own CUDA stream) driven by one host thread
    each: the host<->device copies of one part then overlap the fitting loop of another (the loops themselves do not
    run faster side by side -- tools/stream_overlap_probe.py)."""

    def __init__(self, blends, precision=32, device=None, n_streams=1):
        from .distributed import shard_bounds
        self.blends = list(blends)
        self.precision, self.device = precision, device
        # Dynamic boxes (``resizing=True``, the reference default): every scene is its own sequence of optimiser calls that
        # restart whenever one of its boxes changes (blend.py:99-198), so the scenes of a batch drift apart.  The device keeps
        # per-scene iteration counters and pause flags for that; the batch is then driven as ONE plan (``_fit_dynamic``).
        self.dynamic = any(getattr(c.children[1], "resizing", False) and not c.children[1].parameters[0].fixed
                           for b in self.blends for c in _leaves(b.sources))
        n_streams = 1 if self.dynamic else max(1, min(int(n_streams), len(self.blends)))
        self.parts = [self.blends[a:b] for a, b in shard_bounds(len(self.blends), n_streams)]
        self.plans = [DevicePlan(part, precision=precision, device=device) for part in self.parts]
        self.plan = self.plans[0]
        self.last_transfer_bytes = (0, 0)
        self.replans = 0

    def _each(self, fn):
        if len(self.plans) == 1:
            return [fn(0)]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(self.plans)) as pool:
            return list(pool.map(fn, range(len(self.plans))))

    def upload_observations(self):
        return sum(self._each(lambda i: self.plans[i].upload_observations()))

    def fit(self, max_iter=200, e_rel=1e-3, min_iter=1, noise_factor=0, upload_observations=False, **alg_kwargs):
        check_every = int(alg_kwargs.pop("check_every", 10))
        if alg_kwargs.get("callback") is not None:
            raise NotImplementedError("user callbacks are honoured by Blend.fit (one scene); a batch has no common iteration")
        opts = _fit_options(max_iter, e_rel, min_iter, noise_factor, alg_kwargs, check_every)
        if self.dynamic:
            return self._fit_dynamic(opts, max_iter, e_rel, min_iter, upload_observations)

        def one(i):
            h2d = self._copy_in(i, upload_observations)
            out = self.plans[i].fit(opts)
            return out + (h2d, self._copy_out(i, out))

        return self._finish(self._each(one))

    def _fit_dynamic(self, opts, max_iter, e_rel, min_iter, upload_observations):
        """Batch fit with dynamic boxes: equals ``Blend.fit`` of every scene on its own, bit for bit.

        The device runs all scenes until each one has either converged, spent its budget, failed, or reached an inspection
        point of ITS OWN optimiser call (iterations 10, 20, ... of that call, blend.py:284-292) where it pauses.  The host then
        runs ``src.update()`` on the paused scenes; a scene whose boxes changed restarts (``it = len(loss)``, warm state, its
        ``scheme`` / ``prox_max_iter`` back at the defaults exactly like the reference's restart loop, blend.py:143-152), the
        others resume.  Box changes re-plan the batch (new shapes, new operator tables); state travels through the host."""
        import ctypes
        S = len(self.blends)
        for b in self.blends:
            for src in b.sources:
                src.check_parameters()
        if max_iter <= 0:
            return [(len(b.loss), (-b.loss[-1] if b.loss else None)) for b in self.blends]
        prev_len = np.array([len(b.loss) for b in self.blends], dtype=np.int64)
        it_local = np.zeros(S, dtype=np.int32)
        loss_len = np.zeros(S, dtype=np.int32)
        limit = np.full(S, max_iter, dtype=np.int32)
        prox = np.full(S, opts.prox_max_iter, dtype=np.int32)
        finished = np.zeros(S, dtype=bool)
        loss = np.zeros((S, int(max_iter)))  # this fit's loss histories (host copy; re-uploaded to a re-planned batch)
        opts.max_iter, opts.pause_every = int(max_iter), 10
        opts.check_every = min(opts.check_every, 10)
        h2d = d2h = 0
        lib = nat.lib()
        if self.plan._handle is None:
            self.plan = DevicePlan(self.blends, precision=self.precision, device=self.device)
            self.plans, self.parts = [self.plan], [self.blends]
        elif upload_observations:
            h2d += self.plan.upload_observations()
        plan = self.plan
        # which leaf component (device source index) belongs to which scene / source object
        host_current = True   # the host Parameters hold what the device holds
        need_upload = True    # ... and the device has to be told (first round, or after a re-plan)
        while not finished.all():
            if need_upload:
                h2d += plan.upload_parameters(state=True)
                need_upload = False
            active = (~finished).astype(np.int32)
            nat.check(lib.sb_plan_scene_control(plan._handle, nat.ptr(it_local), nat.ptr(loss_len), nat.ptr(limit), nat.ptr(active),
                                                nat.ptr(prox)))
            launched = ctypes.c_int32()
            nat.check(lib.sb_plan_run(plan._handle, ctypes.byref(opts), int(max_iter) + 1, ctypes.byref(launched)))
            host_current = False
            state = np.zeros(S, dtype=np.int32)
            nat.check(lib.sb_plan_scene_status(plan._handle, nat.ptr(it_local), nat.ptr(loss_len), nat.ptr(state)))
            paused = (state & nat.SCENE_PAUSED) != 0
            # the device reads ImageMorphology.update's rules for every source of a paused scene; only sources it flags (box
            # change, or too close to a threshold to call) need the host -- and only then do parameters travel
            action = plan.inspect() if paused.any() else np.zeros(plan.n_src, dtype=np.int32)
            flagged = [k for k in np.nonzero(action)[0]]
            if flagged or (state & nat.SCENE_FAILED).any():
                d2h += plan.download_parameters(state=True)
                host_current = True
            changed_scene = np.zeros(S, dtype=bool)
            flagged_comps = {id(plan.slots[k]["comp"]) for k in flagged}

            def update(node):
                """``src.update()`` restricted to the leaves the device flagged (the others would return unchanged): a
                multi-component source stops at its first child that changed and re-derives its box (component.py:173-181,
                262-273)."""
                if isinstance(node, CombinedComponent):
                    for child in node.children:
                        try:
                            update(child)
                        except UpdateException:
                            box = node.children[0].bbox.copy()
                            for c in node.children[1:]:
                                box = box | c.bbox
                            node.bbox = box
                            raise
                elif id(node) in flagged_comps:
                    node.update()

            for s in sorted({plan.slots[k]["scene"] for k in flagged}):
                if not paused[s]:
                    continue
                for src in self.blends[s].sources:
                    try:
                        update(src)
                    except UpdateException:
                        changed_scene[s] = True
            for s, b in enumerate(self.blends):
                if finished[s]:
                    continue
                st = int(state[s])
                if st & nat.SCENE_FAILED:
                    nat.check(lib.sb_plan_download_loss(plan._handle, nat.ptr(loss), int(max_iter)))
                    b.loss.extend(loss[s, :loss_len[s]].tolist())
                    for src in b.sources:
                        src.check_parameters()  # raises ArithmeticError naming the parameter (model.py:153-165)
                    raise ArithmeticError("scene %d: a parameter became non-finite during the fit" % s)
                if st & nat.SCENE_PAUSED:
                    if changed_scene[s]:  # blend.py:196-198: restart with it = len(loss); the popped keywords are gone (defaults)
                        it_local[s], prox[s] = 0, 10
                        limit[s] = max_iter - prev_len[s]
                        finished[s] = loss_len[s] >= limit[s]
                    elif st & nat.SCENE_CONV_PENDING:
                        finished[s] = True
                    else:
                        finished[s] = loss_len[s] >= limit[s]
                elif st & (nat.SCENE_CONVERGED | nat.SCENE_EXHAUSTED):
                    finished[s] = True
                elif st == nat.SCENE_RUN and launched.value >= max_iter + 1:
                    finished[s] = True  # cannot happen (budget >= launches); guards an endless loop
            if changed_scene.any():
                plan.replace_sources()  # new boxes / tables; observations stay on the device
                self.replans += 1
                need_upload = True
        if not host_current:
            d2h += plan.download_parameters(state=True)
        nat.check(lib.sb_plan_download_loss(plan._handle, nat.ptr(loss), int(max_iter)))
        d2h += loss.nbytes
        self.last_transfer_bytes = (int(h2d), int(d2h))
        results = []
        for s, b in enumerate(self.blends):
            b.loss.extend(loss[s, :loss_len[s]].tolist())
            results.append((len(b.loss), -b.loss[-1] if b.loss else None))
        return results

    # the three stages of a fit, per plan (BatchPipeline interleaves them across batches)
    def _copy_in(self, i, upload_observations):
        plan = self.plans[i]
        h2d = plan.upload_observations() if upload_observations else 0
        return h2d + plan.upload_parameters(state=True)

    def _copy_out(self, i, out):
        n_iter, loss, status = out
        return self.plans[i].download_parameters(state=True) + n_iter.nbytes + status.nbytes + loss.nbytes

    def _finish(self, outs):
        self.last_transfer_bytes = (int(sum(o[3] for o in outs)), int(sum(o[4] for o in outs)))
        results = []
        for part, (n_iter, loss, status, _, _) in zip(self.parts, outs):
            for s, b in enumerate(part):
                n = int(n_iter[s])
                b.loss.extend(loss[s, :n].tolist())
                if status[s] == nat.SB_ERR_NONFINITE:
                    raise ArithmeticError("scene %d: a parameter became non-finite during the fit" % self.blends.index(b))
                results.append((len(b.loss), -b.loss[-1] if b.loss else None))
        return results

    def close(self):
        for p in self.plans:
            p.close()


class BatchPipeline:
    """Fits a sequence of :class:`BlendBatch` objects back to back on one GPU as a three-stage pipeline: while the device
    runs the loop of batch k (alone, so its kernels see the whole GPU), batch k+1 copies its observations and parameters in
    and batch k-1 copies its results out.  ``depth`` batches are in flight; every batch brings its own device buffers (its
    plans), and a batch object that appears several times in the sequence is fitted again each time, in order.

    ``run`` returns the per-batch results of ``BlendBatch.fit`` in sequence order.  ``prepare(k, batch)`` (optional) runs
    in the worker just before batch k's copy-in (e.g. to load the next set of observations into the host buffers)."""

    def __init__(self, depth=2):
        self.depth = max(1, int(depth))

    def run(self, batches, max_iter=200, e_rel=1e-3, min_iter=1, noise_factor=0, upload_observations=True, prepare=None,
            **alg_kwargs):
        import threading
        import time
        from concurrent.futures import ThreadPoolExecutor
        self.timings = []  # per batch: wall-clock intervals (time.perf_counter) of its three stages
        check_every = int(alg_kwargs.pop("check_every", 10))
        opts = _fit_options(max_iter, e_rel, min_iter, noise_factor, alg_kwargs, check_every)
        batches = list(batches)
        cond = threading.Condition()
        turn = {"in": 0, "loop": 0}
        own = {id(b): threading.Lock() for b in batches}
        failed = []

        def take(stage, k):  # stages are entered in sequence order
            with cond:
                cond.wait_for(lambda: turn[stage] == k or bool(failed))

        def give(stage):
            with cond:
                turn[stage] += 1
                cond.notify_all()

        def work(k):
            b = batches[k]
            idx = range(len(b.plans))
            held = False
            try:
                take("in", k)
                try:
                    if failed:
                        return None
                    own[id(b)].acquire()  # inside the turn: an earlier use of the same batch object finishes first
                    held = True
                    t0 = time.perf_counter()
                    if prepare is not None:
                        prepare(k, b)
                    h2d = [b._copy_in(i, upload_observations) for i in idx]
                    t1 = time.perf_counter()
                finally:
                    give("in")
                take("loop", k)
                try:
                    if failed:
                        return None
                    t2 = time.perf_counter()
                    outs = b._each(lambda i: b.plans[i].fit(opts))
                    t3 = time.perf_counter()
                finally:
                    give("loop")
                res = b._finish([outs[i] + (h2d[i], b._copy_out(i, outs[i])) for i in idx])
                self.timings.append(dict(k=k, copy_in=(t0, t1), loop=(t2, t3), copy_out=(t3, time.perf_counter())))
                return res
            except BaseException as e:  # let the other workers drain instead of waiting for this turn forever
                with cond:
                    failed.append(e)
                    cond.notify_all()
                raise
            finally:
                if held:
                    own[id(b)].release()

        # The worker whose turn it is to start the next device loop must not wait for the interpreter lock behind another
        # worker's Python bookkeeping (CPython hands the lock over every 5 ms by default; the GPU idles meanwhile).
        import sys
        interval = sys.getswitchinterval()
        sys.setswitchinterval(min(interval, 2e-4))
        try:
            with ThreadPoolExecutor(self.depth) as pool:
                return list(pool.map(work, range(len(batches))))
        finally:
            sys.setswitchinterval(interval)
