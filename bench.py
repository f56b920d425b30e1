#!/usr/bin/env python
"""Benchmark of the proximal-gradient fitting path (BASELINE.json metric: PGM iterations/sec per scene).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A "step" is one fit of a batch of independent synthetic scenes for a FIXED number of proximal-gradient iterations
(no early stopping, so that every step does identical work).  One iteration = render + PSF convolution + residual
+ gradient back-propagation + AMSGrad/proximal update of every parameter.

  value  : scene-iterations per second over all GPUs, inputs resident in HBM, CUDA-event timed (max over ranks)
  e2e    : the same through the public API (BatchPipeline over BlendBatch objects) with host buffers: pinned H2D copy of
           the observation cubes, kernel images and parameters, and D2H read-back of fitted parameters, optimiser state
           and losses, every step; the copies of one batch overlap the device loop of the other
  roofline: dominant stage of an iteration, algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline: the oracle restatement of the reference loop on one host core (rank 0, N=1, bounded sample)
Multi-GPU: independent scenes are sharded across ranks (weak scaling, no collective inside the loop); NCCL is used
once per step for the final gather of the packed fitted parameters.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=["cfg2", "cfg3", "cfg4", "cfg4_rot", "cfg5", "tiny"])
    ap.add_argument("--scenes", type=int, default=0, help="scenes per GPU (default per config)")
    ap.add_argument("--iters", type=int, default=0, help="iterations per step (default per config)")
    ap.add_argument("--unique", type=int, default=16, help="distinct synthetic scenes generated per rank (then cycled)")
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--only-headline", action="store_true", help="skip the sub-records of the other BASELINE configurations")
    ap.add_argument("--e2e-mode", default="pipeline", choices=["pipeline", "batch"],
                    help="end-to-end leg: a stream of batches through BatchPipeline (default) or one BlendBatch.fit per step")
    ap.add_argument("--e2e-batches", type=int, default=2, help="--e2e-mode pipeline: BlendBatch objects that take turns (= pipeline depth)")
    ap.add_argument("--e2e-streams", type=int, default=3, help="--e2e-mode batch: plans/streams of the BlendBatch (copy/compute overlap)")
    return ap.parse_args()


DEFAULT_SCENES = {"cfg2": 256, "cfg3": 192, "cfg4": 256, "cfg4_rot": 128, "cfg5": 512, "tiny": 64}
DEFAULT_ITERS = {"cfg2": 50, "cfg3": 50, "cfg4": 50, "cfg4_rot": 50, "cfg5": 50, "tiny": 20}
MULTIRES = ("cfg4", "cfg4_rot")  # cfg4_rot: cfg4 with the low-resolution grid turned by 25 degrees (rotated ResolutionRenderer)


def peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # upper half ~ samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle restatement; the reference itself cannot run here)
# ---------------------------------------------------------------------------------------------------
def _make_scene(config, scene_id):
    from scarlet_b200 import synthetic
    if config == "cfg4_rot":
        return synthetic.make_multires_scene(scene_id, dict(synthetic.CFG4, lr_angle=25.0, config_id=41))
    return synthetic.make_multires_scene(scene_id) if config == "cfg4" else synthetic.make_scene(config, scene_id)


def _make_blend(config, scene, precision=32, device=None):
    from scarlet_b200 import synthetic
    if config in MULTIRES:
        return synthetic.make_multires_blend(scene, precision=precision, device=device)
    return synthetic.make_blend(scene, precision=precision, device=device)


def _config_dict(config):
    from scarlet_b200 import synthetic
    if config in MULTIRES:
        c = synthetic.CFG4
        rot = config == "cfg4_rot"
        return dict(C=8, N=284 if rot else 228, n_ext=c["n_ext"], n_pt=0, psf="gaussian-image", P=c["hr_P"], B=c["B"], symmetric=True,
                    note="5 bands 30x30 at 0.2 arcsec/px (ResolutionRenderer%s) + 3 bands 200x200 at 0.03 arcsec/px (ConvolutionRenderer)"
                         % (", grid rotated by 25 deg" if rot else ""))
    return synthetic.CONFIGS[config]


def _ref_worker(job):
    config, scene_id, iters = job
    from oracle import scenes
    scene = _make_scene(config, scene_id)
    if config in MULTIRES:  # the set-up products of the low-resolution renderer come from the host objects (no GPU involved)
        o = scenes.build_multires_oracle(scene, scenes.multires_setup(_make_blend(config, scene)))
    else:
        o = scenes.build_oracle(scene)
    t0 = time.perf_counter()
    o.fit(max_iter=iters, e_rel=1e-3, min_iter=10 ** 9)
    return time.perf_counter() - t0


def cpu_iters_for(config):
    # iterations of one scene per step of the reference arm (about a second per step and core on the GPU box's host); the
    # one-core cpu_baseline of the default arm runs CPU_BASELINE_FACTOR times as many (about ten seconds)
    return {"cfg2": 120, "cfg3": 40, "cfg4": 4, "cfg4_rot": 3, "cfg5": 100, "tiny": 200}[config]


CPU_BASELINE_FACTOR = 8


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import monotonic_c
    monotonic_c._load()  # in the parent, before the fork: the workers inherit the mapped library (and the driver sees it loaded)
    cores = os.cpu_count() or 1
    iters = cpu_iters_for(args.config)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(k):
            t0 = time.perf_counter()
            pool.map(_ref_worker, [(args.config, 1000 * k + i, iters) for i in range(cores)])
            return time.perf_counter() - t0
        for k in range(args.warmup):
            step(k)
        times = [step(args.warmup + k) for k in range(args.steps)]
    total = float(np.sum(times))
    value = cores * iters * args.steps / total
    sample = ("bounded sample of the workload named in config: %d scenes (one per host thread) x %d iterations per step, %s; "
              "the metric is per scene-iteration" % (cores, iters, args.config))
    from scarlet_b200 import synthetic
    line = {"impl": "reference", "metric": "pgm_scene_iterations_per_sec", "value": value, "unit": "scene-iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.config, _config_dict(args.config), args.scenes or DEFAULT_SCENES[args.config],
                                      args.iters or DEFAULT_ITERS[args.config]),
            "cpu_baseline": {"value": value, "unit": "scene-iterations/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "scene-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(config, cfg, scenes_per_unit, iters):
    """The keys that NAME the workload -- identical in both arms (the reference arm times a bounded sample of it and says so in
    ``cpu_baseline.sample``)."""
    return {"workload": "%s: %d-band %dx%d scene, %d ExtendedSource + %d PointSource, %s PSF %dx%d, box %d, %s"
                        % (config, cfg["C"], cfg["N"], cfg["N"], cfg["n_ext"], cfg["n_pt"], cfg["psf"], cfg["P"], cfg["P"], cfg["B"],
                           "monotonic+symmetry" if cfg["symmetric"] else "monotonic"),
            "scenes_per_gpu": scenes_per_unit, "iterations_per_step": iters, "stop_rule": "disabled (fixed iterations)",
            "l2_policy": "inputs larger than L2 (per-GPU working set reported as device_bytes)"}


# ---------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------
class Env:
    """process-group plumbing shared by the per-config measurements"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            # NCCL prints its version banner on stdout when the communicator is first used: route stdout to stderr until
            # then, so that rank 0's stdout carries exactly one JSON line
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device="cuda:%d" % self.local, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure(env, args, config, S, iters, steps, warmup, headline):
    """Device-resident value, end-to-end value, per-stage profile and roofline of one BASELINE configuration."""
    from scarlet_b200 import BatchPipeline, BlendBatch, _native, synthetic
    torch, dist, rank, world, local = env.torch, env.dist, env.rank, env.world, env.local
    cfg = _config_dict(config)
    uniq = min(S, args.unique)
    base = [_make_scene(config, rank * 100000 + i) for i in range(uniq)]
    blends = [_make_blend(config, base[i % uniq], precision=args.precision, device=local) for i in range(S)]
    batch = BlendBatch(blends, precision=args.precision, device=local)
    plan = batch.plan
    fshape = plan.obs_meta[-1]["metas"][0]["fshape"]
    opts = _native.fit_opts(max_iter=iters, e_rel=1e-3, min_iter=1, prox_max_iter=10, check_every=10 ** 6, fixed_iterations=True)
    init = plan.pack_current()[0]  # initial parameter values (sed, morph, center); the optimiser state starts at zero

    def reset():
        """device-side reset to the initial parameters and a cold optimiser state (outside every timed region)"""
        sed, morph, cen = init
        _native.check(_native.lib().sb_plan_upload_params(plan._handle, 0, _native.ptr(sed), _native.ptr(morph), _native.ptr(cen)))
        _native.check(_native.lib().sb_plan_zero_state(plan._handle))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        plan.sync()

    def gather_results():
        """NCCL gather of the packed fitted parameters (the only collective of the job)."""
        if world == 1:
            return 0
        from scarlet_b200.distributed import gather_device_parameters
        return gather_device_parameters(plan, local)[2]

    # ---- device-resident timing (value): the loop AND the final NCCL gather of the fitted parameters ----------
    def device_step():
        reset()
        barrier()
        t0 = time.perf_counter()
        plan.timer_start()
        plan.fit_enqueue(opts, iters)
        ms = plan.timer_stop()  # CUDA events on the plan's stream around the loop
        t1 = time.perf_counter()
        gather_results()         # NCCL all-gather (N > 1); timed on the host clock between two device syncs
        barrier()
        return ms + (1e3 * (time.perf_counter() - t1) if world > 1 else 0.0)

    for _ in range(warmup):
        device_step()
    sampler = ClockSampler(local) if (headline and rank == 0) else None
    if sampler:
        sampler.start()
    launches0 = plan.kernel_launches
    ms_steps = [device_step() for _ in range(steps)]
    launches = plan.kernel_launches - launches0
    ms_total = env.max_over_ranks(float(np.sum(ms_steps)))
    value = world * S * iters * steps / (ms_total / 1e3)

    # ---- end-to-end through the public API with host buffers (e2e) -------------------------------
    # A stream of batches through BatchPipeline: two BlendBatch objects (own host Parameters, own pinned staging, own device
    # plan) take turns; while the device loops over one, the other copies its results out and the next step's observations,
    # kernel images and parameters in.  Every step's H2D and D2H copies are inside the timed region.
    # (--e2e-mode batch: one BlendBatch.fit per step, split over --e2e-streams plans.)
    h2d = d2h = 0
    pipe_stats = {}
    pipelined = args.e2e_mode == "pipeline"
    extra_batches = []
    if pipelined:
        for _ in range(max(1, args.e2e_batches - 1)):
            blends_b = [_make_blend(config, base[i % uniq], precision=args.precision, device=local) for i in range(S)]
            extra_batches.append(BlendBatch(blends_b, precision=args.precision, device=local))
        turn_batches = [batch] + extra_batches
        batch_e2e = batch
    else:
        batch_e2e = BlendBatch(blends, precision=args.precision, device=local, n_streams=args.e2e_streams) if args.e2e_streams > 1 else batch
        turn_batches = [batch_e2e]
    init_parts = {id(b): [p.pack_current()[0] for p in b.plans] for b in turn_batches}

    def restart(k, b):
        # restore the host Parameters to their initial values and forget the optimiser state (host-side bookkeeping)
        for p, vals in zip(b.plans, init_parts[id(b)]):
            p.forget_state(values=vals)
        for bl in b.blends:
            bl.loss.clear()

    def e2e_steps(n):
        nonlocal h2d, d2h
        if not pipelined:
            total = 0.0
            for _ in range(n):
                restart(0, batch_e2e)
                barrier()
                t0 = time.perf_counter()
                # H2D: data, weights, difference kernels, parameters (pinned staging); the loop; D2H: parameters, state, losses
                batch_e2e.fit(max_iter=iters, e_rel=1e-3, fixed_iterations=True, check_every=10 ** 6, upload_observations=True)
                gather_results()
                barrier()
                total += time.perf_counter() - t0
            h2d, d2h = batch_e2e.last_transfer_bytes
            return total
        for k, b in enumerate(turn_batches):  # first use of each batch: bookkeeping before the clock starts, as in batch mode;
            restart(k, b)                      # later uses restart inside the pipeline (prepare), overlapped with the other loop
        barrier()
        t0 = time.perf_counter()
        pipe = BatchPipeline(depth=len(turn_batches))
        pipe.run([turn_batches[k % len(turn_batches)] for k in range(n)], max_iter=iters, e_rel=1e-3, fixed_iterations=True,
                 check_every=10 ** 6, upload_observations=True,
                 prepare=lambda k, b: restart(k, b) if k >= len(turn_batches) else None)
        starts = sorted(t["loop"][0] for t in pipe.timings)
        if len(starts) > 2:  # period between the starts of consecutive device loops = a step without the pipeline's fill and drain
            pipe_stats["period_ms"] = 1e3 * float(np.median(np.diff(starts)))
            pipe_stats["fill_ms"] = 1e3 * (starts[0] - t0)
        for _ in range(n):
            gather_results()
        barrier()
        dt = time.perf_counter() - t0
        h2d, d2h = batch.last_transfer_bytes
        return dt

    e2e_steps(len(turn_batches) if pipelined else 1)
    e2e_total = env.max_over_ranks(float(e2e_steps(steps)))
    e2e_value = world * S * iters * steps / e2e_total
    clocks = sampler.stop() if sampler else None

    # ---- per-stage CUDA-event profile + roofline (rank 0) -----------------------------------------
    rec = None
    if rank == 0:
        reset()
        plan.profile(opts, 3)
        reset()
        n_prof = 10
        stages = plan.profile(opts, n_prof)
        fused = plan.spectral_mode == 1
        if fused:  # the same marks carry the fused kernels (csrc/scarlet_b200.cu: enqueue_iteration)
            relabel = {"fft_fwd_model": "spec_render", "kmul": "spec_column", "residual_loss": "spec_residual",
                       "kmul_conj": "spec_column_adj", "fft_inv_grad": "spec_grad"}
            # a rotated resampling observation has two more kernels on the marks the cuFFT pipeline uses for its transforms
            extra = {"fft_inv_model": "rot_contract", "fft_fwd_resid": "rot_adjoint"}
            if any(om["metas"][0]["kind"] == 3 for om in plan.obs_meta):
                relabel.update(extra)
            stages = {relabel.get(k, k): v for k, v in stages.items() if k in relabel or k in ("source_update", "advance")}
        peak, peak_src = peaks()
        C, N, B = cfg["C"], cfg["N"], cfg["B"]
        Fy, Fx = fshape
        Fc = Fy * (Fx // 2 + 1)
        eb = 4 if args.precision == 32 else 8
        src_px = cfg["n_ext"] * B * B + cfg["n_pt"] * 81
        Xc = N * (Fx // 2 + 1)  # complex elements of one band's row spectra (Ny rows x Fx/2+1)
        fused_bytes = {  # algorithmic bytes per scene per launch of the fused spectral kernels
            "spec_render": eb * src_px + 2 * eb * C * Xc,
            "spec_column": 2 * 2 * eb * C * Xc + 2 * eb * C * Fc,
            "spec_residual": 2 * 2 * eb * C * Xc + 2 * eb * C * N * N,
            "spec_column_adj": 2 * 2 * eb * C * Xc + 2 * eb * C * Fc,
            "spec_grad": 2 * eb * C * Xc + eb * C * N * N,
        }
        stage_bytes = {  # algorithmic bytes per scene per launch
            "render": eb * (C * N * N + src_px),
            "fft_fwd_model": eb * C * Fy * Fx + 2 * eb * C * Fc,
            "kmul": 3 * 2 * eb * C * Fc,
            "fft_inv_model": eb * C * Fy * Fx + 2 * eb * C * Fc,
            "residual_loss": 4 * eb * C * N * N,
            "fft_fwd_resid": eb * C * Fy * Fx + 2 * eb * C * Fc,
            "kmul_conj": 3 * 2 * eb * C * Fc,
            "fft_inv_grad": eb * C * Fy * Fx + 2 * eb * C * Fc,
            "source_update": eb * src_px * (C + 14),  # SURVEY 8(d) per-source term: B^2 (14 + C)
            "advance": 16,
        }
        stage_bytes.update(fused_bytes)
        if config in MULTIRES:  # two observations on different grids: only the per-source kernel has a closed-form byte count here
            for k in list(stage_bytes):
                if k not in ("source_update", "advance"):
                    stage_bytes[k] = 0
            alg_iter = synthetic.algorithmic_bytes_multires(cfg, eb, frame=tuple(plan.frame_shape), fft_shape=tuple(fshape))
        else:
            alg_iter = synthetic.algorithmic_bytes(cfg, fshape, eb)
        own = {k: v for k, v in stages.items() if not k.startswith("fft_") and stage_bytes.get(k, 0) > 0}
        # the kernel furthest below its roofline (lowest algorithmic GB/s) among those that matter (>= 5 % of the iteration)
        iter_ms = sum(stages.values())
        gbs = {k: stage_bytes[k] * S / (v / 1e3) / 1e9 for k, v in own.items() if v > 0}
        cand = [k for k in gbs if stages[k] >= 0.05 * iter_ms] or list(gbs)
        dom = min(cand, key=lambda k: gbs[k])
        dom_ms = stages[dom]
        achieved = gbs[dom]
        whole = alg_iter * S / (ms_total / steps / iters / 1e3) / 1e9
        # DRAM bytes of the same kernel from the committed `ncu --set full` capture of this workload (per launch)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_%s_s%d.json" % (config, S))
        kname = {"source_update": "k_update_warp", "spec_render": "k_spec_render", "spec_column": "k_spec_column",
                 "spec_column_adj": "k_spec_column", "spec_residual": "k_spec_residual", "spec_grad": "k_spec_grad"}.get(dom)
        if args.precision == 32 and fused and kname and os.path.exists(tpath):
            try:
                traffic = float(json.load(open(tpath))["kernels"][kname]["dram_bytes_per_launch"])
                traffic_src = "profiles/" + os.path.basename(tpath)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": stage_bytes[dom] * S,
                    "kernel_ms": dom_ms, "kernel_share_of_iteration": dom_ms / iter_ms,
                    "iteration": {"algorithmic_bytes_per_scene": alg_iter, "achieved": whole, "frac": whole / peak,
                                  "note": "SURVEY.md 8(d) whole-iteration accounting (cuFFT-pipeline byte model), device-resident value"},
                    "stages_ms": stages,
                    "stages_gbs": {k: (stage_bytes[k] * S / (v / 1e3) / 1e9 if v > 0 and stage_bytes.get(k, 0) > 0 else None)
                                   for k, v in stages.items()}}

        if "rot_contract" in stages:  # dense FP32 contraction over the half plane (not tensor-core work: float32 parity)
            H, W = plan.obs_meta[0]["metas"][0]["shape"][1:]
            Kh, lrC = Fy * (Fx // 2 + 1), plan.obs_meta[0]["metas"][0]["shape"][0]
            flop = {"rot_contract": lrC * Kh * (4 * H * W + 6 * H), "rot_adjoint": lrC * Kh * (4 * H * W + 8 * H + 6)}
            mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
            fp32_peak = 148 * 128 * 2 * mhz * 1e6 / 1e12
            roofline["fp32_contraction"] = {k: {"flop_per_scene": v, "ms": stages[k], "achieved_tflops": v * S / (stages[k] / 1e3) / 1e12,
                                                "frac_of_fp32_peak": v * S / (stages[k] / 1e3) / 1e12 / fp32_peak} for k, v in flop.items()}
            roofline["fp32_contraction"]["fp32_peak_tflops"] = fp32_peak

        # single-scene latency (the 200 it/s target of the north star is a per-scene figure)
        single = None
        if headline and not args.no_single:
            one = BlendBatch([_make_blend(config, base[0], precision=args.precision, device=local)], precision=args.precision, device=local)
            o1 = _native.fit_opts(max_iter=200, e_rel=1e-3, fixed_iterations=True, check_every=10 ** 6)
            for _ in range(3):
                one.plan.forget_state()
                one.plan.upload_parameters(state=True)
                one.plan.timer_start()
                one.plan.fit_enqueue(o1, 200)
                ms1 = one.plan.timer_stop()
            single = 200 / (ms1 / 1e3)
            one.plan.close()

        details = {"e2e": ("BatchPipeline(depth=%d): %d BlendBatch objects take turns, the copies of one overlap the loop of another"
                           % (len(turn_batches), len(turn_batches))
                           if pipelined else "one BlendBatch.fit per step over %d plans/streams" % len(batch_e2e.plans)),
                   "fft_grid": list(fshape), "unique_scenes_per_gpu": uniq, "device_bytes_per_gpu": plan.device_bytes,
                   "single_scene_iterations_per_sec": single,
                   "spectral": "fused row/column kernels" if fused else "cuFFT",
                   "cufft_execs_per_iteration": 0 if fused else 4 * len(plan.obs_meta),
                   "kernels_per_iteration": launches // max(steps * iters, 1), "precision": args.precision,
                   "per_scene_psf": config not in MULTIRES,
                   "value_includes": "the proximal-gradient loop (CUDA events) + the final NCCL gather of fitted parameters (N > 1)"}
        if pipe_stats.get("period_ms"):
            details["e2e_pipeline"] = {"steady_state_ms_per_step": pipe_stats["period_ms"], "fill_ms": pipe_stats["fill_ms"],
                                       "steady_state_value": world * S * iters / (pipe_stats["period_ms"] / 1e3),
                                       "note": "e2e.value times exactly K steps, i.e. it carries the first copy-in and the last copy-out "
                                               "un-overlapped; the period between consecutive loop starts is what a long stream sees"}
        if cfg.get("note"):
            details["observations"] = cfg["note"]
        rec = {"value": value, "unit": "scene-iterations/s", "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
               "config": workload_config(config, cfg, S, iters), "details": details, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "scene-iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "ms_per_step": 1e3 * e2e_total / steps},
               "gpu_launches": int(launches), "roofline": roofline}
    barrier()
    if batch_e2e is not batch:
        batch_e2e.close()
    for b in extra_batches:
        b.close()
    plan.close()
    return rec


def measure_dynamic(env, args, S=64, iters=50, start_box=31):
    """cfg2-shaped scenes with dynamic boxes ON (``resizing=True``, the reference default): one ``BlendBatch.fit`` through the
    public API with host buffers; every scene inspects its sources every 10 iterations of its own optimiser call, boxes that
    changed re-plan the batch (scarlet_b200/blend.py:_fit_dynamic).  The boxes start at 31 x 31 around sources that want 41 x 41
    and more, so the early inspections do resize.  Host-driven, hence end-to-end only."""
    from scarlet_b200 import BlendBatch, synthetic
    cfg = dict(synthetic.CONFIGS["cfg2"], B=start_box, resizing=True, config_id=22)
    uniq = min(S, args.unique)
    base = [synthetic.make_scene(cfg, env.rank * 100000 + i) for i in range(uniq)]

    def run():
        blends = [synthetic.make_blend(base[i % uniq], precision=args.precision, device=env.local) for i in range(S)]
        batch = BlendBatch(blends, precision=args.precision, device=env.local)
        env.torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = batch.fit(max_iter=iters, e_rel=1e-9, upload_observations=True)
        env.torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out = (dt, sum(r[0] for r in res), batch.replans, batch.last_transfer_bytes,
               sorted(set(int(src.parameters[1].shape[0]) for b in blends for src in b.sources)))
        batch.close()
        return out

    run()
    runs = sorted((run() for _ in range(3)), key=lambda r: r[0])  # host-driven and short: the median of three fits
    dt, n_it, replans, (h2d, d2h), sizes = runs[1]
    spread = [runs[0][0], runs[2][0]]
    dt = env.max_over_ranks(dt)
    return {"value": None, "unit": "scene-iterations/s",
            "config": workload_config("cfg2 with dynamic boxes (resizing=True, start %dx%d)" % (start_box, start_box), cfg, S, iters),
            "e2e": {"value": env.world * n_it / dt, "unit": "scene-iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * dt},
            "replans": replans, "final_box_sizes": sizes, "fit_seconds_min_max": spread,
            "note": "host-driven inspection + re-plan; no device-resident figure"}


def run_b200(args):
    from scarlet_b200 import _native
    if _native.lib().sb_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- scarlet_b200 has no CPU fallback")
    env = Env()
    S = args.scenes or DEFAULT_SCENES[args.config]
    iters = args.iters or DEFAULT_ITERS[args.config]
    head = measure(env, args, args.config, S, iters, args.steps, args.warmup, headline=True)
    # the other BASELINE configurations, a few steps each, as sub-records of the same line (device value, e2e, roofline)
    others = {}
    if args.config == "cfg3" and not args.only_headline:
        for c in ("cfg5", "cfg2", "cfg4", "cfg4_rot"):
            try:  # a sub-record must never take the headline down with it
                rec = measure(env, args, c, DEFAULT_SCENES[c], DEFAULT_ITERS[c], min(args.steps, 5), 3, headline=False)
            except Exception as e:  # noqa: BLE001 -- reported in the line
                rec = None
                if env.rank == 0:
                    others[c] = {"error": "%s: %s" % (type(e).__name__, e)}
            if rec is not None:
                others[c] = {"value": rec["value"], "unit": rec["unit"], "ms_per_step": rec["ms_per_step"], "steps": rec["steps"],
                             "config": rec["config"], "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"],
                             "roofline": {"iteration": rec["roofline"]["iteration"], "kernel": rec["roofline"]["kernel"],
                                          "frac": rec["roofline"]["frac"], "stages_ms": rec["roofline"]["stages_ms"],
                                          "fp32_contraction": rec["roofline"].get("fp32_contraction")},
                             "fft_grid": rec["details"]["fft_grid"], "device_bytes_per_gpu": rec["details"]["device_bytes_per_gpu"]}
    if args.config == "cfg3" and not args.only_headline:
        # boxes too small at the start: every source resizes in the first rounds (worst case); then BASELINE's 200 iterations,
        # over which the re-plans of the first rounds amortise
        for name, kw in (("cfg2_dynamic", {}), ("cfg2_dynamic_200", dict(S=256, iters=200, start_box=41))):
            try:
                dyn = measure_dynamic(env, args, **kw)
            except Exception as e:  # noqa: BLE001
                dyn = {"error": "%s: %s" % (type(e).__name__, e)}
            if env.rank == 0:
                others[name] = dyn
    line = None
    if env.rank == 0:
        cpu = None
        if env.world == 1 and not args.no_cpu_baseline:
            from oracle import monotonic_c
            monotonic_c._load()
            n_cpu = CPU_BASELINE_FACTOR * cpu_iters_for(args.config)
            _ref_worker((args.config, 0, 2))
            dt = _ref_worker((args.config, 0, n_cpu))
            cpu = {"value": n_cpu / dt, "unit": "scene-iterations/s", "cores": 1, "kind": "port",
                   "sample": "1 %s scene x %d iterations on one host core (oracle restatement, NumPy f64 + C sweep); host has %d cores"
                             % (args.config, n_cpu, os.cpu_count() or 1)}
        line = {"metric": "pgm_scene_iterations_per_sec", "value": head["value"], "unit": head["unit"], "n_gpus": env.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
                "config": head["config"], "details": head["details"], "clocks": head["clocks"], "e2e": head["e2e"],
                "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "cpu_baseline": cpu, "configs": others}
    if env.world > 1:
        env.dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
