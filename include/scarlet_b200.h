/* scarlet_b200 -- C ABI of the B200-native proximal-gradient fitting path.
 *
 * This is the drop-in boundary for ONE hot path of pmelchior/scarlet: the per-iteration work of
 * `Blend.fit` (reference scarlet/blend.py:85-198).  Every entry point names the reference
 * interface it replaces.  Conventions (all entry points):
 *   - plain pointers and sizes, no C++ / torch types; host pointers are borrowed for the call only
 *     (C-contiguous), the plan owns every device allocation, the cuFFT plans and one CUDA stream;
 *   - return value: 0 on success, negative on error; sb_last_error() returns a message for the
 *     calling thread; no exception ever crosses this boundary;
 *   - one plan <-> one device <-> one stream; a plan is not thread-safe;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails (-2).
 */
#ifndef SCARLET_B200_H
#define SCARLET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_ARG (-1)
#define SB_ERR_CUDA (-2)
#define SB_ERR_CUFFT (-3)
#define SB_ERR_NONFINITE (-4) /* a parameter became inf/nan: scarlet/model.py:153-165 -> ArithmeticError */

/* ---- constraint op-codes: the device image of the Constraint plugin surface -------------------
 * reference scarlet/constraint.py: ConstraintChain 58-80, PositivityConstraint 83-92,
 * NormalizationConstraint 95-114, MonotonicityConstraint 183-234 (use_mask=False,
 * fit_center_radius=0), SymmetryConstraint 262-273, CenterOnConstraint 276-287. */
enum {
    SB_OP_MONOTONIC = 1, /* iarg = index into the monotonic-operator tables, farg = min_gradient */
    SB_OP_SYMMETRY = 2,  /* farg = strength */
    SB_OP_POSITIVITY = 3, /* farg = zero */
    SB_OP_CENTER_ON = 4,  /* farg = tiny */
    SB_OP_NORMALIZE = 5   /* iarg = 0 sum, 1 max */
};
#define SB_MAX_CHAIN_OPS 8
#define SB_MAX_CHANNELS 16
#define SB_MAX_OBS 4

typedef struct sb_op {
    int32_t code;
    int32_t iarg;
    double farg;
} sb_op;

typedef struct sb_chain_desc { /* one ConstraintChain (constraint.py:58-80) */
    int32_t n_ops;
    int32_t repeat;
    sb_op ops[SB_MAX_CHAIN_OPS];
} sb_chain_desc;

/* One radial-monotonicity operator = the arguments of the reference's native
 * prox_weighted_monotonic (scarlet/operators_pybind11.cc:14-36, built by operator.py:62-96). */
typedef struct sb_mono_desc {
    int32_t n_pix;          /* By*Bx */
    int32_t n_off;          /* 8 */
    int32_t n_idx;          /* len(dist_idx), normally n_pix-1 */
    int32_t _pad;
    const double *weights;  /* [n_off][n_pix], row-major */
    const int32_t *offsets; /* [n_off] flat neighbour offsets */
    const int32_t *dist_idx; /* [n_idx] pixels in sweep order */
} sb_mono_desc;

/* One Observation matched to the model frame (observation.py:59-114, renderer.py:164-202). */
typedef struct sb_obs_desc {
    int32_t kind;     /* 0 = ConvolutionRenderer (fft), 1 = NullRenderer, 2 = ResolutionRenderer (aligned grids; H x W is the
                         low-resolution cube, Fy x Fx the reference's _fft_shape, renderer.py:288-290), 3 = ResolutionRenderer with
                         rotated grids (renderer.py:318-363) */
    int32_t C, H, W;  /* data cube */
    int32_t chan_off; /* first model-frame channel (channel map = slice, renderer.py:26-51) */
    int32_t oy, ox;   /* position of data pixel (0,0) in the model frame (translation only) */
    int32_t Fy, Fx;   /* FFT grid (fft.py:116-167) */
    int32_t khat_shared; /* 1: one K^ for all scenes, 0: one per scene */
    int32_t psf_shift;   /* kind 0 only: ConvolutionRenderer(psf_shift=...) (renderer.py:172-177, 220-227): the difference kernel
                            is moved by a fitted (dy, dx) with fft.shift before every convolution; the parameter travels in the
                            centre arrays behind the sources' entries: observation-major, one slot per scene */
    int32_t shift_Fy, shift_Fx; /* the reference's fast grid of that shift: _get_fft_shape(kernel, kernel, padding=10) */
    int32_t shift_fixed;
    double shift_step;   /* 1e-2 (renderer.py:176) */
} sb_obs_desc;

/* One FactorizedComponent (component.py:119-193). */
typedef struct sb_source_desc {
    int32_t kind;    /* 0 = image morphology (ExtendedSource), 1 = PointSourceMorphology on a GaussianPSF */
    int32_t By, Bx;  /* morphology box */
    int32_t oy, ox;  /* box origin in the model frame (may be negative / overhang, bbox.py:279-301) */
    int32_t chain;   /* morphology constraint chain index, -1 = none */
    int32_t sed_chain; /* spectrum constraint chain index, -1 = none */
    int32_t sed_is_f32; /* spectrum Parameter is float32 on the host (rounded after each update) */
    int32_t morph_fixed, sed_fixed;
    int32_t shifting;       /* kind 0 only: the model uses fft.shift(image, shift) (morphology.py:124-130, fft.py:399-428); the
                               free ``shift`` parameter then travels in the centre arrays like a point-source centre */
    int32_t shift_Fy, shift_Fx; /* the reference's fast grid of that shift: _get_fft_shape(image, image, padding=10) */
    int32_t resizing;       /* kind 0 only: the box adapts during the fit (ImageMorphology(resizing=True), morphology.py:132-207);
                               sb_plan_inspect evaluates the shrink / grow rules for such sources */
    double shift_step;      /* constant step of the shift parameter (1e-1, morphology.py:672-675) */
    double morph_step;      /* constant step of the image / center parameter */
    double sed_step_factor; /* relative_step factor (parameter.py:126-129); <0: constant step = sed_step_min[0] */
    double sed_step_min[SB_MAX_CHANNELS]; /* per-band minimum step (spectrum.py:56, source.py:412-416) */
} sb_source_desc;

typedef struct sb_batch_desc {
    int32_t precision; /* 32 or 64: real type of grids, morphologies and optimiser state on the device */
    int32_t n_scenes;
    int32_t C, Ny, Nx; /* model frame (identical for all scenes of a batch) */
    int32_t n_obs;
    sb_obs_desc obs[SB_MAX_OBS];
    int32_t n_sources; /* over all scenes */
    int32_t n_chains;
    int32_t n_mono;
    int32_t psf_boxsize;                /* model GaussianPSF box (point sources), 0 if none */
    double psf_sigma[SB_MAX_CHANNELS];  /* model GaussianPSF sigma per band */
    const int32_t *scene_src_start;     /* [n_scenes+1] */
    const sb_source_desc *sources;      /* [n_sources] */
    const sb_chain_desc *chains;        /* [n_chains] */
    const sb_mono_desc *mono;           /* [n_mono] */
} sb_batch_desc;

typedef struct sb_fit_opts { /* Blend.fit / proxmin.adaprox arguments, blend.py:85,165-180 */
    int32_t max_iter;
    int32_t min_iter;
    int32_t prox_max_iter;
    int32_t check_every; /* host polls the device stop flags every this many iterations */
    int32_t fixed_iterations; /* 1: ignore the stop rule (benchmark mode) */
    int32_t overwrite_vhat_at_it0; /* oracle switch (1), SURVEY 8c open point (1) */
    int32_t resume;    /* 1: continue the previous sb_plan_fit call of this plan (iteration counter, stop flags and loss
                          history are kept) -- one proxmin.adaprox call split at the points where Blend._callback
                          inspects the sources (every 10 iterations, blend.py:284-292) */
    int32_t run_until; /* stop this call when the iteration counter reaches run_until (0: max_iter) */
    double e_rel;
    double b1, b2, eps;
    int32_t pause_every; /* > 0: a scene whose adaprox iteration counter is a positive multiple of this pauses after that
                            iteration (state SB_SCENE_PAUSED) so that the host can inspect its sources -- Blend._callback's
                            src.update() every 10 iterations (blend.py:284-292).  0: never */
    int32_t _pad1;
} sb_fit_opts;

/* per-scene run state (sb_plan_scene_status); distinct bits */
#define SB_SCENE_RUN 0
#define SB_SCENE_CONVERGED 1   /* the stop rule fired (blend.py:294-299) */
#define SB_SCENE_PAUSED 2      /* inspection point reached; | SB_SCENE_CONV_PENDING if the stop rule also fired there */
#define SB_SCENE_EXHAUSTED 4   /* loss history reached the scene's limit (max_iter) */
#define SB_SCENE_FAILED 8      /* non-finite parameter */
#define SB_SCENE_CONV_PENDING 16

typedef struct sb_plan sb_plan;

const char *sb_last_error(void);
int sb_device_count(void);
const char *sb_version(void);
/* hash of the sources the library was compiled from (scarlet_b200/_build.py:source_hash): the loader refuses a binary
 * whose hash differs from the sources next to it instead of calling a stale ABI */
const char *sb_source_hash(void);

/* Transform lengths of the fused spectral kernels (csrc/spectral.cuh): smallest supported length >= need, 0 if the
 * request exceeds the largest one (the plan then runs the convolutions through cuFFT on any 2/3/5/7-smooth grid). */
int sb_fft_supported_length(int need);

/* ---- plan life cycle: replaces the closures Blend.fit hands to proxmin.adaprox (blend.py:103-180) */
int sb_plan_create(const sb_batch_desc *desc, int device, sb_plan **out);
void sb_plan_destroy(sb_plan *plan);
int64_t sb_plan_device_bytes(const sb_plan *plan);
/* 1: the convolutions of this plan run in the fused row/column spectral kernels; 0: cuFFT + separate kernels
 * (grids with unsupported lengths, NullRenderer observations, or SB_SPECTRAL=cufft in the environment). */
int sb_plan_spectral_mode(const sb_plan *plan);
/* Diagnostic: histogram of the proximal sub-iterations (1..prox_max_iter, blend.py:145) the warp and grouped update kernels ran
 * per source since the last call.  enable=1 starts (or continues) counting and zeroes the counters after reading;
 * enable=0 reads and switches the counters off.  out16 may be NULL. */
int sb_plan_prox_histogram(sb_plan *plan, int enable, int64_t *out16);

/* ---- batches whose scenes restart independently (dynamic boxes: every scene is its own sequence of proxmin.adaprox calls,
 * blend.py:99-198).  Each scene carries its own iteration counter of the running call (proxmin's ``it``), the length of its
 * loss history in this fit, an iteration budget and a run flag.  All arrays have n_scenes entries; NULL leaves a table as is.
 *   it_local  : counter of the running adaprox call (0 after a restart)
 *   loss_len  : entries of the loss history written so far (the next loss goes to column loss_len)
 *   limit     : the scene stops (SB_SCENE_EXHAUSTED) when loss_len reaches limit
 *   active    : 1 = run, 0 = skip (also clears the error flags)
 *   prox_iter : per-scene prox_max_iter (a restarted call runs with the default 10, blend.py:143-145)
 * sb_plan_run launches iterations (no reset of anything) until no scene is running or max_launches is reached; the number
 * of running scenes is polled every opts->check_every launches. */
int sb_plan_scene_control(sb_plan *plan, const int32_t *it_local, const int32_t *loss_len, const int32_t *limit,
                          const int32_t *active, const int32_t *prox_iter);
int sb_plan_scene_status(sb_plan *plan, int32_t *it_local, int32_t *loss_len, int32_t *state);
int sb_plan_run(sb_plan *plan, const sb_fit_opts *opts, int max_launches, int32_t *launched);
/* Dynamic boxes: for every source of a PAUSED scene that is marked `resizing`, evaluate ImageMorphology.update's rules on the
 * device (morphology.py:52-68, 132-207): action[k] = new box size (shrink: outer rings entirely <= 0; grow: the next gradient
 * update pulls more than 0.1 of the peak towards an edge), 0 = keep, -1 = too close to a threshold to call (the host decides).
 * action has n_sources entries.  The host only has to look at sources with action != 0. */
int sb_plan_inspect(sb_plan *plan, int32_t *action);
/* Replace the sources of a plan (same scenes, frame and observations; new boxes / chains / tables): everything on the
 * observation side -- data, weights, K^, spectral buffers, tensor maps -- stays where it is.  Parameters and optimiser state
 * are NOT carried over: upload them afterwards.  desc->obs must equal the plan's. */
int sb_plan_set_sources(sb_plan *plan, const sb_batch_desc *desc);
/* loss histories [n_scenes][n_cols] (n_cols <= the capacity set by the largest max_iter seen so far) */
int sb_plan_download_loss(sb_plan *plan, double *loss, int n_cols);
/* ... and back (a re-planned batch continues its histories: the stop rule compares with the previous entry) */
int sb_plan_upload_loss(sb_plan *plan, const double *loss, int n_cols);

/* Observation data: data/weights float32 [n_scenes][C][H][W] (frame dtype, frame.py:29); K^ = rfftn of the
 * padded, ifftshifted difference kernel (renderer.py:198-202, fft.py:255-273) as interleaved complex128
 * [n_scenes or 1][C][Fy][Fx/2+1]; loss_const[n_scenes] = log_norm (observation.py:172-186) plus the
 * chi^2 of data pixels outside the model frame. */
int sb_plan_upload_observation(sb_plan *plan, int obs, const float *data, const float *weights,
                               const double *khat, const double *loss_const);
/* The same with float64 data / weights (a model frame of dtype float64 makes Observation.match keep float64 cubes,
 * observation.py:75-84): a precision-64 plan stores them as they are, a precision-32 plan rounds on the device. */
int sb_plan_upload_observation_f64(sb_plan *plan, int obs, const double *data, const double *weights,
                                   const double *khat, const double *loss_const);
/* Difference-kernel images instead of a host-computed K^ (what fft.convolve does on every call, fft.py:385-388):
 * kernels float64 [n][C][Py][Px] with n = 1 for an observation declared khat_shared, else n_scenes; (y0, x0) = grid
 * index of kernel pixel (0,0) before wrapping, i.e. the reference's centre-pad + ifftshift placement
 * (fft.py:82-113, 255-273; -(P//2) for odd P).  The device pads, wraps, transforms in double precision and stores
 * K^/(Fy Fx) in the plan's precision.  The grid must satisfy F >= N + max(-y0, P-1+y0) per axis (no wrap-around
 * inside the frame), else SB_ERR_ARG. */
int sb_plan_upload_kernels(sb_plan *plan, int obs, const double *kernels, int Py, int Px, int y0, int x0);
/* Resampling observation (kind 2; ResolutionRenderer, renderer.py:262-547): K^ = rfft2 of the centre-padded difference
 * kernel WITHOUT origin shift goes through sb_plan_upload_observation(khat); this call adds the Fourier-shift matrices
 * Ey complex128 [H][Fy], Ex complex128 [W][Fx/2+1] (exp(-2 pi i f s), Nyquist bins real, for the low-resolution pixel
 * rows / columns in model-frame grid coordinates) and h2 = (pixel-scale ratio)^2. */
int sb_plan_upload_resampling(sb_plan *plan, int obs, const double *ey, const double *ex, double h2);
/* Rotated resampling observation (kind 3; the rotated branch of ResolutionRenderer, renderer.py:318-363, 498-524: `shifts` per
 * low-resolution row, `other_shifts` per column, sinc_shift along both axes).  K^ as for kind 2; a complex128 [H][Fy][Fx/2+1]
 * and b complex128 [W][Fy][Fx/2+1] are the half-plane multipliers of the two-axis Fourier shifts attached to the rows and the
 * columns (Hermitian part of the phase ramps, the model's entering conjugated and carrying the translation to the grid
 * origin); LR[c,i,j] = h2 sum_kx w_kx Re sum_ky K^ conj(M^) a_i b_j.  H, W <= 32. */
int sb_plan_upload_resampling_rot(sb_plan *plan, int obs, const double *a, const double *b, double h2);
/* Pinned host memory for staging buffers: copies from/to it are asynchronous DMA. */
void *sb_host_alloc(int64_t bytes);
void sb_host_free(void *p);
/* Pack / unpack the values of many small host arrays (the Parameter objects of a batch) into / from one
 * contiguous float64 buffer: src[i] points at count[i] contiguous elements, float32 if is_f32[i] else float64. */
int sb_host_gather_f64(double *dst, const void *const *src, const int64_t *count, const int32_t *is_f32, int64_t n);
int sb_host_scatter_f64(const double *src, void *const *dst, const int64_t *count, const int32_t *is_f32, int64_t n);

/* Parameters and optimiser state, concatenated over sources in plan order (Parameter.m/v/vhat,
 * parameter.py:42-71; warm start blend.py:154-163).  sed arrays: [n_sources][C]; morph arrays: concatenated
 * By*Bx images of the kind-0 sources; center arrays: [n_point + n_shifting][2] in source order (point-source centres and
 * the shifts of shifting image morphologies).  which: 0 = value, 1 = m, 2 = v, 3 = vhat.
 * Any pointer may be NULL (skipped). */
int sb_plan_upload_params(sb_plan *plan, int which, const double *sed, const double *morph, const double *center);
int sb_plan_download_params(sb_plan *plan, int which, double *sed, double *morph, double *center);
/* m = v = vhat = 0 for every parameter: the cold start of blend.py:154-163 without shipping zeros. */
int sb_plan_zero_state(sb_plan *plan);

/* One gradient evaluation at the current parameters without an update: model (blend.py:200-244),
 * rendered models (observation.py:131-145), loss (blend.py:259-274) and the gradient of the loss wrt every
 * parameter (what autograd.grad returns at blend.py:118).  Any output may be NULL.
 * model: float32/64 per plan precision cast to double [n_scenes][C][Ny][Nx]; rendered: [n_scenes][C_o][H][W]
 * of observation `obs`; loss: [n_scenes]; gradients laid out like the parameters. */
int sb_plan_evaluate(sb_plan *plan, int obs, double *model, double *rendered, double *loss,
                     double *g_sed, double *g_morph, double *g_center);

/* The fitting loop: replaces proxmin.adaprox(X, grad, step, prox=..., scheme="amsgrad", ...) + Blend._callback
 * (blend.py:165-180, 276-302).  n_iter_out[n_scenes]: gradient evaluations per scene (= len(blend.loss) growth);
 * loss_out[n_scenes][max_iter] (entries >= n_iter are untouched); status_out[n_scenes]: 0 ok, SB_ERR_NONFINITE. */
int sb_plan_fit(sb_plan *plan, const sb_fit_opts *opts, int32_t *n_iter_out, double *loss_out, int32_t *status_out);
/* Asynchronous halves of sb_plan_fit for callers that own the timing (bench): enqueue n iterations on the plan's
 * stream starting at iteration counter it0 without any host synchronisation, then wait. */
int sb_plan_fit_enqueue(sb_plan *plan, const sb_fit_opts *opts, int n_iterations);
int sb_plan_sync(sb_plan *plan);
/* CUDA-event timing of the work enqueued on the plan's stream between the two calls (milliseconds). */
int sb_plan_timer_start(sb_plan *plan);
int sb_plan_timer_stop(sb_plan *plan, float *ms);
/* per-kernel-class device time accumulated by sb_plan_profile_iterations (events around every stage) */
#define SB_N_STAGES 10
int sb_plan_profile_iterations(sb_plan *plan, const sb_fit_opts *opts, int n_iterations, float *stage_ms /*[SB_N_STAGES]*/);
const char *sb_stage_name(int stage);
int64_t sb_plan_kernel_launches(const sb_plan *plan); /* launches of this library's kernels + cuFFT execs so far */
void *sb_plan_stream(sb_plan *plan);                   /* cudaStream_t */
/* device pointers of the packed fitted parameters (for the NCCL gather of results): sed, morph, center */
int sb_plan_device_params(sb_plan *plan, void **sed, int64_t *n_sed, void **morph, int64_t *n_morph, int *elem_bytes);

/* ---- single-operator entry points (host buffers in, result in place / out) -----------------------------
 * sb_monotonic_*: same argument list as the reference binding prox_weighted_monotonic(flat_img, weights,
 * offsets, dist_idx, min_gradient) (operators_pybind11.cc:14-36, bound at :243-246), executed as a wavefront
 * kernel; n_img images of n_pix pixels are processed with the same operator. */
int sb_monotonic_f32(float *flat_img, const float *weights, const int32_t *offsets, int n_off,
                     const int32_t *dist_idx, int n_idx, int n_pix, float min_gradient, int n_img, int device);
int sb_monotonic_f64(double *flat_img, const double *weights, const int32_t *offsets, int n_off,
                     const int32_t *dist_idx, int n_idx, int n_pix, double min_gradient, int n_img, int device);
/* A whole ConstraintChain on n_img images of shape (By,Bx) (constraint.py:76-80). */
int sb_prox_chain_f32(float *img, int By, int Bx, int n_img, const sb_chain_desc *chain,
                      const sb_mono_desc *mono, int n_mono, int device);
int sb_prox_chain_f64(double *img, int By, int Bx, int n_img, const sb_chain_desc *chain,
                      const sb_mono_desc *mono, int n_mono, int device);
/* fft.convolve(Fourier(image), kernel, axes=(1,2)) with a precomputed K^ (fft.py:368-396); adjoint=1 applies
 * conj(K^) (the VJP).  image/out: [C][Ny][Nx]; khat: complex128 [C][Fy][Fx/2+1]. */
int sb_fft_convolve_f32(const float *image, int C, int Ny, int Nx, const double *khat, int Fy, int Fx,
                        int adjoint, float *out, int device);
int sb_fft_convolve_f64(const double *image, int C, int Ny, int Nx, const double *khat, int Fy, int Fx,
                        int adjoint, double *out, int device);

/* Same argument list as the reference binding apply_filter(image, values, y_start, y_end, x_start, x_end, result)
 * (operators_pybind11.cc:39-56, bound at :248-249): result = sum_n values[n] * image shifted by tap n, zero outside the image;
 * taps are accumulated per pixel in order n = 0..n_taps-1 with separate multiply and add, like the reference's loop. */
int sb_apply_filter_f32(const float *image, int H, int W, const float *values, const int32_t *y_start, const int32_t *y_end,
                        const int32_t *x_start, const int32_t *x_end, int n_taps, float *result, int device);
int sb_apply_filter_f64(const double *image, int H, int W, const double *values, const int32_t *y_start, const int32_t *y_end,
                        const int32_t *x_start, const int32_t *x_end, int n_taps, double *result, int device);

#ifdef __cplusplus
}
#endif
#endif /* SCARLET_B200_H */
