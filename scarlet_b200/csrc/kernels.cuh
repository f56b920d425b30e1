// Device kernels of the scarlet_b200 fitting path (sm_100a).  Templated on the real type T (float = product
// path, double = high-precision twin used to separate algorithmic from rounding differences in the parity tests).
#pragma once
#include "common.cuh"

namespace sb {

// ======================================================================================================
// radial monotonicity: wavefront sweep
// reference: scarlet/operators_pybind11.cc:14-36 (sequential sweep in distance order).  A pixel only reads
// neighbours that are STRICTLY closer to the centre (operator.py:612-614), so pixels fall into dependency
// levels; all pixels of a level are independent.  Per pixel the arithmetic (sum over the positive-weight
// neighbours in offset order, un-fused multiply/add, times (1-min_gradient), min) is the reference's, hence
// the result is bit-identical to the sequential sweep.
// ======================================================================================================
template <typename T> struct __align__(16) W4 { T a, b, c, d; };

template <typename T, int NB> struct MonoRec {
    int pix;
    uint2 nbr[NB / 4];
    W4<T> w[NB / 4];
};

template <typename T, int NB>
__device__ __forceinline__ void mono_load(MonoRec<T, NB> &r, const DevMono &mo, int j) {
    r.pix = __ldg(mo.pix + j);
    const uint2 *nbr = reinterpret_cast<const uint2 *>(mo.code);
    const W4<T> *w = reinterpret_cast<const W4<T> *>(mo.w);
#pragma unroll
    for (int g = 0; g < NB / 4; ++g) {
        r.nbr[g] = nbr[(size_t)g * mo.n_tasks + j];
        r.w[g] = w[(size_t)g * mo.n_tasks + j];
    }
}

template <typename T, int NB>
__device__ __forceinline__ void mono_apply(T *img, const MonoRec<T, NB> &r, T keep) {
    T ref = T(0);
#pragma unroll
    for (int g = 0; g < NB / 4; ++g) {
        const unsigned n0 = r.nbr[g].x & 0xffffu, n1 = r.nbr[g].x >> 16, n2 = r.nbr[g].y & 0xffffu, n3 = r.nbr[g].y >> 16;
        if (n0 != 0xffffu) ref = add_rn(ref, mul_rn(img[n0], r.w[g].a));
        if (n1 != 0xffffu) ref = add_rn(ref, mul_rn(img[n1], r.w[g].b));
        if (n2 != 0xffffu) ref = add_rn(ref, mul_rn(img[n2], r.w[g].c));
        if (n3 != 0xffffu) ref = add_rn(ref, mul_rn(img[n3], r.w[g].d));
    }
    const T cap = mul_rn(ref, keep);
    if (cap < img[r.pix]) img[r.pix] = cap;
}

// img lives in shared memory; every thread of the block calls this; ends with a block barrier.
template <typename T, int NB> __device__ void mono_sweep(T *img, const DevMono &mo, T min_gradient) {
    const T keep = T(1) - min_gradient;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int *ls = mo.level_start;
    MonoRec<T, NB> nxt;
    int beg = __ldg(ls), end = __ldg(ls + 1);
    bool have = beg + tid < end;
    if (have) mono_load<T, NB>(nxt, mo, beg + tid);
    for (int L = 0; L < mo.n_levels; ++L) {
        const MonoRec<T, NB> cur = nxt;
        const bool chave = have;
        const int cbeg = beg, cend = end;
        if (L + 1 < mo.n_levels) { // prefetch this thread's first task of the next level before the barrier
            beg = cend;
            end = __ldg(ls + L + 2);
            have = beg + tid < end;
            if (have) mono_load<T, NB>(nxt, mo, beg + tid);
        }
        if (chave) mono_apply<T, NB>(img, cur, keep);
        for (int j = cbeg + tid + nt; j < cend; j += nt) { // levels wider than the block (rare)
            MonoRec<T, NB> extra;
            mono_load<T, NB>(extra, mo, j);
            mono_apply<T, NB>(img, extra, keep);
        }
        __syncthreads();
    }
}

template <typename T> __device__ __forceinline__ void mono_sweep_any(T *img, const DevMono &mo, T min_gradient) {
    if (mo.nb == 4)
        mono_sweep<T, 4>(img, mo, min_gradient);
    else
        mono_sweep<T, 8>(img, mo, min_gradient);
}

// ======================================================================================================
// ConstraintChain on one image in shared memory (constraint.py:58-114, 183-287; operator.py:274-293)
// ======================================================================================================
template <typename T>
__device__ void apply_chain(T *a, int By, int Bx, const DevChain &ch, const DevMono *monos, double *red) {
    const int n = By * Bx, tid = threadIdx.x, nt = blockDim.x;
    for (int r = 0; r < ch.repeat; ++r) {
        for (int o = 0; o < ch.n_ops; ++o) {
            const sb_op op = ch.ops[o];
            switch (op.code) {
            case SB_OP_MONOTONIC:
                mono_sweep_any<T>(a, monos[op.iarg], (T)op.farg);
                break;
            case SB_OP_SYMMETRY: { // blend with the 180-degree rotation; even axes are extended by one zero line
                const T s = (T)op.farg, hs = (T)(0.5 * op.farg), om = (T)(1.0 - op.farg);
                const int Hy = By + ((By & 1) == 0), Wx = Bx + ((Bx & 1) == 0);
                for (int p = tid; p < n; p += nt) {
                    const int y = p / Bx, x = p - y * Bx;
                    const int yr = Hy - 1 - y, xr = Wx - 1 - x;
                    if (yr >= By || xr >= Bx) {
                        const T u = a[p];
                        a[p] = hs * (u + T(0)) + om * u;
                    } else {
                        const int q = yr * Bx + xr;
                        if (q > p) {
                            const T u = a[p], v = a[q];
                            a[p] = hs * (u + v) + om * u;
                            a[q] = hs * (v + u) + om * v;
                        } else if (q == p) {
                            const T u = a[p];
                            a[p] = hs * (u + u) + om * u;
                        }
                    }
                }
                (void)s;
                __syncthreads();
                break;
            }
            case SB_OP_POSITIVITY: {
                const T zero = (T)op.farg;
                for (int p = tid; p < n; p += nt) a[p] = a[p] < zero ? zero : a[p]; // np.maximum: NaN propagates
                __syncthreads();
                break;
            }
            case SB_OP_CENTER_ON: {
                if (tid == 0) {
                    const int c = (By / 2) * Bx + Bx / 2;
                    const T tiny = (T)op.farg;
                    a[c] = a[c] < tiny ? tiny : a[c]; // Python max(a, tiny): a NaN stays
                }
                __syncthreads();
                break;
            }
            case SB_OP_NORMALIZE: {
                double acc;
                if (op.iarg == 1) {
                    acc = -INFINITY;
                    for (int p = tid; p < n; p += nt) acc = fmax(acc, (double)a[p]);
                    acc = block_max(acc, red);
                } else {
                    acc = 0.0;
                    for (int p = tid; p < n; p += nt) acc += (double)a[p];
                    acc = block_sum(acc, red);
                }
                const T den = (T)acc;
                for (int p = tid; p < n; p += nt) a[p] = a[p] / den;
                __syncthreads();
                break;
            }
            default:
                break;
            }
        }
    }
}

// element-wise subset of the chain for 1-D spectra, executed by ONE thread (C <= 16 values)
__device__ inline void apply_chain_1d(double *a, int n, const DevChain &ch) {
    for (int r = 0; r < ch.repeat; ++r)
        for (int o = 0; o < ch.n_ops; ++o) {
            const sb_op op = ch.ops[o];
            if (op.code == SB_OP_POSITIVITY) {
                for (int i = 0; i < n; ++i) a[i] = a[i] < op.farg ? op.farg : a[i]; // NaN propagates like np.maximum
            } else if (op.code == SB_OP_NORMALIZE) {
                double acc = op.iarg == 1 ? -INFINITY : 0.0;
                for (int i = 0; i < n; ++i) acc = op.iarg == 1 ? fmax(acc, a[i]) : acc + a[i];
                for (int i = 0; i < n; ++i) a[i] /= acc;
            }
        }
}

// ======================================================================================================
// K1  render: model[c,y,x] = sum_k sed_k[c] * morph_k[y-oy_k, x-ox_k]   (gather form, deterministic order)
// reference: component.py:144-171 (outer product), blend.py:17-27, 200-244 (insertion in source order).
// The model is written straight into the zero-padded real FFT grid of every observation (image at the grid
// origin -- the reference's centring + ifftshift cancel against fftshift + centre-crop).
// ======================================================================================================
template <typename T> struct RenderArgs {
    const DevSource *src;
    const int *scene_src_start;
    const double *sed; // [n_src][C]
    const T *morph;
    const T *pmorph;
    const T *smorph; // Fourier-shifted images of the shifting sources
    int C, Ny, Nx, n_obs;
    DevObs<T> obs[SB_MAX_OBS];
    const int *done;
    T *model_out; // optional [S][C][Ny][Nx]
};

template <typename T> __global__ void __launch_bounds__(256) k_render(const RenderArgs<T> a) {
    const int s = blockIdx.z;
    if (a.done[s]) return;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.Nx || y >= a.Ny) return;
    T acc[SB_MAXC];
#pragma unroll
    for (int c = 0; c < SB_MAXC; ++c) acc[c] = T(0);
    const int k0 = a.scene_src_start[s], k1 = a.scene_src_start[s + 1];
    const int C = a.C;
    for (int k = k0; k < k1; ++k) {
        const DevSource &d = a.src[k];
        const int by = y - d.oy, bx = x - d.ox;
        if ((unsigned)by < (unsigned)d.By && (unsigned)bx < (unsigned)d.Bx) {
            const double *sed = a.sed + (size_t)k * C;
            if (d.kind == 0) {
                const T mv = (d.shifting ? a.smorph : a.morph)[d.morph_off + (size_t)by * d.Bx + bx];
#pragma unroll
                for (int c = 0; c < SB_MAXC; ++c)
                    if (c < C) acc[c] += (T)sed[c] * mv;
            } else {
                const T *pm = a.pmorph + d.morph_off + (size_t)by * d.Bx + bx;
                const int plane = d.By * d.Bx;
#pragma unroll
                for (int c = 0; c < SB_MAXC; ++c)
                    if (c < C) acc[c] += (T)sed[c] * pm[(size_t)c * plane];
            }
        }
    }
    for (int o = 0; o < a.n_obs; ++o) {
        const DevObs<T> &ob = a.obs[o];
#pragma unroll
        for (int c = 0; c < SB_MAXC; ++c) {
            const int co = c - ob.chan_off;
            if (c < C && co >= 0 && co < ob.C) ob.A[(((size_t)s * ob.C + co) * ob.Fy + y) * ob.Fx + x] = acc[c];
        }
    }
    if (a.model_out) {
#pragma unroll
        for (int c = 0; c < SB_MAXC; ++c)
            if (c < C) a.model_out[(((size_t)s * C + c) * a.Ny + y) * a.Nx + x] = acc[c];
    }
}

// ======================================================================================================
// K2  k-space product  X^ *= K^  (or conj K^ for the adjoint)      reference: fft.py:316-331, 385-396
// ======================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
k_kmul(typename Cx<T>::type *__restrict__ X, const typename Cx<T>::type *__restrict__ K, long long per_scene,
       long long total, int shared, int conj, const int *done) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / per_scene;
        if (done && done[s]) continue;
        const typename Cx<T>::type k = K[shared ? i - s * per_scene : i];
        const typename Cx<T>::type v = X[i];
        const T ki = conj ? -k.y : k.y;
        typename Cx<T>::type r;
        r.x = v.x * k.x - v.y * ki;
        r.y = v.x * ki + v.y * k.x;
        X[i] = r;
    }
}

// ======================================================================================================
// K3  residual + loss: r = w (render - data), written into the (zero-padded) gradient grid; loss partials
// reference: observation.py:147-170, renderer.py:130-161 (match_shape and its VJP)
// ======================================================================================================
template <typename T> struct ResidualArgs {
    DevObs<T> ob;
    int Ny, Nx;
    const int *done;
    double *partials; // [S*C][gridDim.y*gridDim.x] chi^2 partial sums of this observation (deterministic reduction)
    T *rendered_out;  // optional [S][C][H][W]
};

template <typename T> __global__ void __launch_bounds__(256) k_residual(const ResidualArgs<T> a) {
    __shared__ double red[40];
    const DevObs<T> &ob = a.ob;
    const int sc = blockIdx.z; // scene * C + channel
    const int s = sc / ob.C;
    if (a.done[s]) return;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    double part = 0.0;
    if (x < a.Nx && y < a.Ny) {
        const int dy = y - ob.oy, dx = x - ob.ox;
        T r = T(0);
        if ((unsigned)dy < (unsigned)ob.H && (unsigned)dx < (unsigned)ob.W) {
            const size_t di = ((size_t)sc * ob.H + dy) * ob.W + dx;
            const T m = ob.B[((size_t)sc * ob.Fy + y) * ob.Fx + x];
            const T w = ob.weights[di];
            const T diff = m - ob.data[di];
            r = w * diff;
            part = (double)w * (double)diff * (double)diff;
            if (a.rendered_out) a.rendered_out[di] = m;
        }
        ob.A[((size_t)sc * ob.Fy + y) * ob.Fx + x] = r;
    }
    // block reduction (blockDim = 32x8 -> linear thread id)
    const int lin = threadIdx.y * 32 + threadIdx.x;
    part = warp_sum(part);
    if ((lin & 31) == 0) red[lin >> 5] = part;
    __syncthreads();
    if (lin < 32) {
        part = lin < 8 ? red[lin] : 0.0;
        part = warp_sum(part);
        if (lin == 0) a.partials[(size_t)sc * (gridDim.x * gridDim.y) + blockIdx.y * gridDim.x + blockIdx.x] = part;
    }
}

// ======================================================================================================
// pixel-integrated Gaussian (psf.py:129-142) and its derivative
// ======================================================================================================
__device__ __forceinline__ double gauss_int(double x, double sigma) {
    const double s2 = sqrt(2.0) * sigma;
    return sqrt(M_PI / 2) * sigma * (1 - erfc((0.5 - x) / s2) + 1 - erfc((2 * x + 1) / (2 * s2)));
}
__device__ __forceinline__ double gauss_int_deriv(double x, double sigma) {
    return exp(-((x + 0.5) * (x + 0.5)) / (2 * sigma * sigma)) - exp(-((x - 0.5) * (x - 0.5)) / (2 * sigma * sigma));
}

// numpy's float mean for short vectors (pairwise_sum: n < 8 sequential; else 8 accumulators), parameter.py:126-129
template <typename TS> __device__ inline TS numpy_mean(const double *x, int n) {
    TS res;
    if (n < 8) {
        res = TS(0);
        for (int i = 0; i < n; ++i) res += (TS)x[i];
    } else {
        TS r[8];
        for (int j = 0; j < 8; ++j) r[j] = (TS)x[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += (TS)x[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += (TS)x[i];
    }
    return res / (TS)n;
}

// ======================================================================================================
// K4/K5  per-source backward + AMSGrad adaprox update + constraint projections
// reference: gradients = what autograd.grad yields at blend.py:118 (explicit forms lite/models.py:206-216);
// update = proxmin.adaprox(scheme="amsgrad", prox_max_iter) as called at blend.py:165-180 (structure mirrored
// in-tree at lite/parameters.py:274-305); step sizes blend.py:135-138, parameter.py:126-129.
// ======================================================================================================
template <typename T> struct UpdateArgs {
    const DevSource *src;
    int C, Ny, Nx, n_obs;
    DevObs<T> obs[SB_MAX_OBS];
    double *sed, *sed_m, *sed_v, *sed_vhat;
    T *morph, *morph_m, *morph_v, *morph_vhat;
    double *center, *cen_m, *cen_v, *cen_vhat;
    T *pmorph;
    const DevChain *chains;
    const DevMono *monos;
    const int *it_ptr; // [S] per-scene iteration counter of the running adaprox call (proxmin's ``it``)
    const int *prox_iter; // optional [S]: per-scene prox_max_iter (a restarted scene falls back to the default, blend.py:143-145)
    const int *done;
    int *status;
    FitScalars fs;
    int psf_b;
    double psf_sigma[SB_MAXC];
    int npix_max;   // shared-memory image length (largest box, padded)
    int npix_shift; // ... of the largest box of a shifting source (0: none): three more scratch images
    int mode;     // 0 = update, 1 = gradients only
    double *g_sed, *g_morph, *g_center;
    T *smorph;              // packed like morph: Fourier-shifted images of the shifting sources (what the model uses)
    T *toep;                // [n_shift][8][toep_len]: Re/Im Toeplitz vectors of the shift along y and x and their d/ds
    int toep_len;           // 2*Bmax-1
    const int *shift_list;  // k_shift_apply: source index per CTA
    const int *work;        // generic kernel: source index per block (NULL: block index)
    // grouped fast path (k_update_fast): one 64-thread group per source, groups of a CTA share one operator table
    const int *fast_groups; // [n_cta][fast_G] source index or -1
    int fast_G, fast_npix;  // groups per CTA, shared-memory image length per group
    int fast_table_cap;     // task capacity of the shared-memory table
    T *scratch_x, *scratch_ps; // packed like the morphologies: gradient-step result x and metric psi
    unsigned long long *prox_hist; // optional [16]: histogram of proximal sub-iterations executed per source (diagnostic)
};

// gradient of the loss wrt the model at frame pixel (y,x), channel c: sum over the observations that see c
template <typename T>
__device__ __forceinline__ double grad_at(const UpdateArgs<T> &a, int s, int c, int y, int x) {
    double g = 0.0;
    for (int o = 0; o < a.n_obs; ++o) {
        const DevObs<T> &ob = a.obs[o];
        const int co = c - ob.chan_off;
        if (co >= 0 && co < ob.C) g += (double)ob.B[(((size_t)s * ob.C + co) * ob.Bh + y) * ob.Bw + x];
    }
    return g;
}

// The C band gradients at one pixel.  With a single observation the loads are issued back to back (straight-line code, no
// loop over observations between them) so that one memory round trip serves all bands; otherwise band by band via grad_at.
template <typename T, int CMAX>
__device__ __forceinline__ void grad_bands(const UpdateArgs<T> &a, int s, int C, int y, int x, T (&v)[CMAX]) {
    const DevObs<T> &ob = a.obs[0];
    const size_t plane = (size_t)ob.Bh * ob.Bw;
    const T *b = ob.B + ((size_t)s * ob.C * ob.Bh + y) * ob.Bw + x;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        const int co = c - ob.chan_off;
        v[c] = (c < C && co >= 0 && co < ob.C) ? b[(size_t)co * plane] : T(0);
    }
}

// The same for several observations: per observation the band loads are issued back to back, the sums over the observations
// that see a band are formed in double in observation order (exactly grad_at's arithmetic, one memory round trip per
// observation instead of one per band).
template <typename T, int CMAX>
__device__ __forceinline__ void grad_bands_multi(const UpdateArgs<T> &a, int s, int C, int y, int x, T (&v)[CMAX]) {
    double acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.0;
    for (int o = 0; o < a.n_obs; ++o) {
        const DevObs<T> &ob = a.obs[o];
        const size_t plane = (size_t)ob.Bh * ob.Bw;
        const T *b = ob.B + ((size_t)s * ob.C * ob.Bh + y) * ob.Bw + x;
        T t[CMAX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            const int co = c - ob.chan_off;
            t[c] = (c < C && co >= 0 && co < ob.C) ? b[(size_t)co * plane] : T(0);
        }
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            const int co = c - ob.chan_off;
            if (c < C && co >= 0 && co < ob.C) acc[c] += (double)t[c];
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) v[c] = (T)acc[c];
}

template <typename T> __device__ __forceinline__ int prox_max_iter_of(const UpdateArgs<T> &a, int scene) {
    return a.prox_iter ? a.prox_iter[scene] : a.fs.prox_max_iter;
}

// AMSGrad moments (Reddi, Kale & Kumar 2018, no bias correction) -- proxmin's _amsgrad_phi_psi as restated
// in oracle/scarlet_oracle.py:amsgrad_phi_psi.  Returns psi; m, v, vhat updated in place.
__device__ __forceinline__ double amsgrad(double g, double &m, double &v, double &vhat, int it, const FitScalars &fs) {
    m = (1 - fs.b1) * g + fs.b1 * m;
    v = (1 - fs.b2) * (g * g) + fs.b2 * v;
    vhat = (it == 0 && fs.overwrite_vhat_at_it0) ? v : fmax(vhat, v);
    return sqrt(fs.eps > 0 ? fmax(vhat, fs.eps) : vhat);
}

// spectrum update, executed by one thread.  g: gradient [C].
template <typename T>
__device__ void sed_update(const UpdateArgs<T> &a, const DevSource &d, int k, const double *g, int it) {
    const int C = a.C;
    double *x = a.sed + (size_t)k * C, *m = a.sed_m + (size_t)k * C, *v = a.sed_v + (size_t)k * C,
           *vh = a.sed_vhat + (size_t)k * C;
    double alpha[SB_MAXC], psi[SB_MAXC], xn[SB_MAXC], z[SB_MAXC], zn[SB_MAXC];
    double rel = 0.0;
    if (d.sed_step_factor >= 0) {
        if (d.sed_is_f32)
            rel = (double)((float)d.sed_step_factor * numpy_mean<float>(x, C));
        else
            rel = d.sed_step_factor * numpy_mean<double>(x, C);
    }
    double psimax = 0.0;
    for (int c = 0; c < C; ++c) {
        alpha[c] = d.sed_step_factor >= 0 ? fmax(d.sed_step_min[c], rel) : d.sed_step_min[0];
        double mm = m[c], vv = v[c], vvh = vh[c];
        psi[c] = amsgrad(g[c], mm, vv, vvh, it, a.fs);
        m[c] = mm, v[c] = vv, vh[c] = vvh;
        xn[c] = x[c] - alpha[c] * mm / psi[c];
        if (d.sed_is_f32) xn[c] = (double)(float)xn[c];
        psimax = fmax(psimax, psi[c]);
    }
    if (d.sed_chain >= 0) {
        const DevChain &ch = a.chains[d.sed_chain];
        for (int c = 0; c < C; ++c) z[c] = xn[c];
        for (int sub = 0; sub < prox_max_iter_of(a, d.scene); ++sub) {
            for (int c = 0; c < C; ++c) {
                const double gamma = alpha[c] / psimax;
                zn[c] = z[c] - gamma / alpha[c] * psi[c] * (z[c] - xn[c]);
            }
            apply_chain_1d(zn, C, ch);
            double dd = 0.0, nn = 0.0;
            for (int c = 0; c < C; ++c) {
                dd += (zn[c] - z[c]) * (zn[c] - z[c]);
                nn += z[c] * z[c];
                z[c] = zn[c];
            }
            if (dd <= a.fs.e_rel * a.fs.e_rel * nn) break;
        }
        for (int c = 0; c < C; ++c) xn[c] = d.sed_is_f32 ? (double)(float)z[c] : z[c];
    }
    bool bad = false;
    for (int c = 0; c < C; ++c) {
        x[c] = xn[c];
        bad |= !isfinite(xn[c]);
    }
    if (bad) atomicExch(a.status + d.scene, SB_ERR_NONFINITE);
}

// normalised point-source morphology planes for the current centre (morphology.py:503-507, psf.py:103-127)
// fy/fx: shared scratch of >= 16 doubles each.  Every thread of the block must call this.
template <typename T>
__device__ void point_planes(const UpdateArgs<T> &a, const DevSource &d, double cy, double cx, double *fy, double *fx) {
    const int b = d.By, tid = threadIdx.x, nt = blockDim.x, C = a.C;
    const double offy = cy - (d.oy + b / 2.0), offx = cx - (d.ox + b / 2.0);
    T *pm = a.pmorph + d.morph_off;
    for (int c = 0; c < C; ++c) {
        const double sg = a.psf_sigma[c];
        if (tid < b)
            fy[tid] = gauss_int((double)(tid - b / 2) - offy, sg);
        else if (tid < 2 * b)
            fx[tid - b] = gauss_int((double)(tid - b - b / 2) - offx, sg);
        __syncthreads();
        double Sy = 0.0, Sx = 0.0;
        for (int j = 0; j < b; ++j) Sy += fy[j], Sx += fx[j];
        for (int p = tid; p < b * b; p += nt) {
            const int by = p / b, bx = p - by * b;
            pm[(size_t)c * b * b + p] = (T)((fy[by] * fx[bx]) / (Sy * Sx));
        }
        __syncthreads();
    }
}

template <typename T> __global__ void __launch_bounds__(128) k_point_morph(const UpdateArgs<T> a, int n_src) {
    __shared__ double fy[16], fx[16];
    const int k = blockIdx.x;
    if (k >= n_src) return;
    const DevSource &d = a.src[k];
    if (d.kind != 1) return;
    point_planes<T>(a, d, a.center[2 * d.point_idx], a.center[2 * d.point_idx + 1], fy, fx);
}

template <typename T> __device__ void update_point(const UpdateArgs<T> &a, const DevSource &d, int k, double *red) {
    __shared__ double fy[16], fx[16], dfy[16], dfx[16], gsed[SB_MAXC];
    const int b = d.By, tid = threadIdx.x, nt = blockDim.x, C = a.C, s = d.scene, it = a.it_ptr[d.scene];
    double *cen = a.center + 2 * d.point_idx;
    const double cy = cen[0], cx = cen[1];
    const double offy = cy - (d.oy + b / 2.0), offx = cx - (d.ox + b / 2.0);
    const double *sed = a.sed + (size_t)k * C;
    double gc0 = 0.0, gc1 = 0.0;
    for (int c = 0; c < C; ++c) {
        const double sg = a.psf_sigma[c];
        if (tid < b) {
            const double X = (double)(tid - b / 2) - offy;
            fy[tid] = gauss_int(X, sg);
            dfy[tid] = -gauss_int_deriv(X, sg);
        } else if (tid < 2 * b) {
            const double X = (double)(tid - b - b / 2) - offx;
            fx[tid - b] = gauss_int(X, sg);
            dfx[tid - b] = -gauss_int_deriv(X, sg);
        }
        __syncthreads();
        double Sy = 0.0, Sx = 0.0, Dy = 0.0, Dx = 0.0;
        for (int j = 0; j < b; ++j) Sy += fy[j], Sx += fx[j], Dy += dfy[j], Dx += dfx[j];
        double gs = 0.0;
        const double sc = sed[c];
        for (int p = tid; p < b * b; p += nt) {
            const int by = p / b, bx = p - by * b, y = d.oy + by, x = d.ox + bx;
            if ((unsigned)y < (unsigned)a.Ny && (unsigned)x < (unsigned)a.Nx) {
                const double g = grad_at<T>(a, s, c, y, x);
                const double ny = fy[by] / Sy, nx = fx[bx] / Sx;
                const double dny = dfy[by] / Sy - fy[by] * Dy / (Sy * Sy), dnx = dfx[bx] / Sx - fx[bx] * Dx / (Sx * Sx);
                gs += g * (ny * nx);
                const double Gm = sc * g;
                gc0 += Gm * (dny * nx);
                gc1 += Gm * (ny * dnx);
            }
        }
        gs = block_sum(gs, red);
        if (tid == 0) gsed[c] = gs;
        __syncthreads();
    }
    gc0 = block_sum(gc0, red);
    gc1 = block_sum(gc1, red);
    if (a.mode == 1) {
        if (tid == 0) {
            if (a.g_sed)
                for (int c = 0; c < C; ++c) a.g_sed[(size_t)k * C + c] = gsed[c];
            if (a.g_center) a.g_center[2 * d.point_idx] = gc0, a.g_center[2 * d.point_idx + 1] = gc1;
        }
        return;
    }
    if (tid == 0) {
        if (!d.morph_fixed) {
            const double g2[2] = {gc0, gc1};
            double *m = a.cen_m + 2 * d.point_idx, *v = a.cen_v + 2 * d.point_idx, *vh = a.cen_vhat + 2 * d.point_idx;
            for (int i = 0; i < 2; ++i) {
                double mm = m[i], vv = v[i], vvh = vh[i];
                const double psi = amsgrad(g2[i], mm, vv, vvh, it, a.fs);
                m[i] = mm, v[i] = vv, vh[i] = vvh;
                cen[i] = cen[i] - d.morph_step * mm / psi;
                if (!isfinite(cen[i])) atomicExch(a.status + s, SB_ERR_NONFINITE);
            }
        }
        if (!d.sed_fixed) sed_update<T>(a, d, k, gsed, it);
    }
    __syncthreads();
    __threadfence_block();
    point_planes<T>(a, d, cen[0], cen[1], fy, fx);
}

// ======================================================================================================
// Sub-pixel shifted image morphologies: ImageMorphology(shifting=True) -> fft.shift (morphology.py:124-130, fft.py:399-428)
//
// The reference pads the image to a fast grid F, multiplies its rfftn by exp(-2 pi i (fftfreq_y s0 + rfftfreq_x s1)) and
// transforms back.  Restricted to the B x B box this is EXACTLY   o = Re(Cy) u Re(Tx)^T - Im(Cy) u Im(Tx)^T   with the
// Toeplitz vectors  Cy[d] = 1/Fy sum_k exp(2 pi i (k d - m_k s0)/Fy)  (m_k: signed frequency, complex inverse transform
// along y) and  Tx[d] = 1/Fx sum_{k<=Fx/2} c_k exp(2 pi i k (d - s1)/Fx)  (c = 1,2,...,2,1: what the real inverse transform
// along x does) -- including the non-Hermitian Nyquist row a fractional shift creates (checked against the reference's
// fft.shift to 2e-15).  The vectors and their derivatives wrt the shift are rebuilt for the current shift by direct
// summation in double precision (any F, no FFT), the products are small dense contractions in shared memory.
// ======================================================================================================
template <typename T> __device__ __forceinline__ T *toep_vec(const UpdateArgs<T> &a, const DevSource &d, int v) {
    return a.toep + ((size_t)d.toep_off * 8 + v) * a.toep_len;
}

// vectors (0..7) = Re Cy, Im Cy, d/ds0 Re Cy, d/ds0 Im Cy, Re Tx, Im Tx, d/ds1 Re Tx, d/ds1 Im Tx; entry j <-> offset d = j-(B-1)
template <typename T> __device__ void shift_vectors(const UpdateArgs<T> &a, const DevSource &d, double s0, double s1) {
    const int By = d.By, Bx = d.Bx, ny = 2 * By - 1, nx = 2 * Bx - 1;
    for (int idx = threadIdx.x; idx < ny + nx; idx += blockDim.x) {
        const bool ydir = idx < ny;
        const int F = ydir ? d.shift_Fy : d.shift_Fx, off = ydir ? idx - (By - 1) : idx - ny - (Bx - 1);
        const double sh = ydir ? s0 : s1, w = 2.0 * M_PI / F;
        double re = 0, im = 0, dre = 0, dim = 0;
        if (ydir) {
            for (int k = 0; k < F; ++k) {
                const int m = k < (F + 1) / 2 ? k : k - F; // numpy.fft.fftfreq
                double sn, cs;
                sincos(w * ((double)k * off - (double)m * sh), &sn, &cs);
                re += cs, im += sn;
                const double f = -w * m; // d/ds0 of the phase
                dre += -f * sn, dim += f * cs;
            }
        } else {
            for (int k = 0; k <= F / 2; ++k) {
                const double c = (k == 0 || 2 * k == F) ? 1.0 : 2.0;
                double sn, cs;
                sincos(w * k * ((double)off - sh), &sn, &cs);
                re += c * cs, im += c * sn;
                const double f = -w * k;
                dre += -c * f * sn, dim += c * f * cs;
            }
        }
        const int j = ydir ? idx : idx - ny, base = ydir ? 0 : 4;
        toep_vec<T>(a, d, base + 0)[j] = (T)(re / F), toep_vec<T>(a, d, base + 1)[j] = (T)(im / F);
        toep_vec<T>(a, d, base + 2)[j] = (T)(dre / F), toep_vec<T>(a, d, base + 3)[j] = (T)(dim / F);
    }
}

// out[y',x'] = sum_{y,x} (Ay[y'-y] in[y,x] Bx[x'-x]) as two passes through the scratch image w (all in shared memory)
// transpose = true applies the transposed operators (Ay[y-y'], Bx[x-x']): the vector-Jacobian product
template <typename T, bool TRANSPOSE>
__device__ void toeplitz_apply(const T *in, T *w, T *out, const T *Ay, const T *Bxv, int By, int Bx, bool accumulate, T sign) {
    const int n = By * Bx;
    for (int p = threadIdx.x; p < n; p += blockDim.x) { // along x
        const int y = p / Bx, x1 = p - y * Bx;
        T acc = T(0);
        for (int x = 0; x < Bx; ++x) acc += in[y * Bx + x] * Bxv[(TRANSPOSE ? x - x1 : x1 - x) + Bx - 1];
        w[p] = acc;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) { // along y
        const int y1 = p / Bx, x = p - y1 * Bx;
        T acc = T(0);
        for (int y = 0; y < By; ++y) acc += w[y * Bx + x] * Ay[(TRANSPOSE ? y - y1 : y1 - y) + By - 1];
        out[p] = accumulate ? out[p] + sign * acc : sign * acc;
    }
    __syncthreads();
}

// one CTA per shifting source: Toeplitz vectors for the current shift, then smorph = shift(morph)
template <typename T> __global__ void __launch_bounds__(128) k_shift_apply(const UpdateArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int k = a.shift_list[blockIdx.x];
    const DevSource &d = a.src[k];
    if (a.done[d.scene]) return;
    const int n = d.By * d.Bx;
    T *u = reinterpret_cast<T *>(smem), *w = u + a.npix_shift, *o = w + a.npix_shift;
    shift_vectors<T>(a, d, a.center[2 * d.point_idx], a.center[2 * d.point_idx + 1]);
    for (int p = threadIdx.x; p < n; p += blockDim.x) u[p] = a.morph[d.morph_off + p];
    __threadfence_block();
    __syncthreads();
    toeplitz_apply<T, false>(u, w, o, toep_vec<T>(a, d, 0), toep_vec<T>(a, d, 4), d.By, d.Bx, false, T(1));
    toeplitz_apply<T, false>(u, w, o, toep_vec<T>(a, d, 1), toep_vec<T>(a, d, 5), d.By, d.Bx, true, T(-1));
    for (int p = threadIdx.x; p < n; p += blockDim.x) a.smorph[d.morph_off + p] = o[p];
}

// Generic per-source update (one CTA per source): any box that fits one image in shared memory, any constraint chain.
// Only the image being projected (zn) lives in shared memory -- the sweep and the symmetry partner need random access;
// the gradient-step result x, the metric psi and the running iterate z (= the morphology array itself) are streamed
// through global memory, each element always by the same thread.  Shifting sources additionally use three scratch
// images in shared memory for the Toeplitz products.
template <typename T> __device__ void update_extended(const UpdateArgs<T> &a, const DevSource &d, int k, unsigned char *smem) {
    const int tid = threadIdx.x, nt = blockDim.x, C = a.C, s = d.scene, it = a.it_ptr[d.scene];
    const int n = d.By * d.Bx, Bx = d.Bx;
    T *zn = reinterpret_cast<T *>(smem);
    double *red = reinterpret_cast<double *>(zn + a.npix_max);
    double *gsum = red + 40;
    T *t0 = reinterpret_cast<T *>(gsum + SB_MAXC); // shifting sources only: three more images (a.npix_shift each)
    T *t1 = t0 + a.npix_shift, *t2 = t1 + a.npix_shift;

    double sedv[SB_MAXC], gs[SB_MAXC];
#pragma unroll
    for (int c = 0; c < SB_MAXC; ++c) {
        sedv[c] = c < C ? a.sed[(size_t)k * C + c] : 0.0;
        gs[c] = 0.0;
    }
    T *mp = a.morph + d.morph_off, *mm = a.morph_m + d.morph_off, *mv = a.morph_v + d.morph_off,
      *mvh = a.morph_vhat + d.morph_off, *gx = a.scratch_x + d.morph_off, *gp = a.scratch_ps + d.morph_off;
    const double alpha = d.morph_step;
    const bool upd = a.mode == 0 && !d.morph_fixed;
    double pmax = 0.0;
    double gshift0 = 0.0, gshift1 = 0.0;
    if (d.shifting) {
        // the model holds the SHIFTED image: gather the gradient wrt it (zn), the spectrum gradient against it, then pull the
        // gradient back to the image (t0) and to the shift through the transposed Toeplitz operators
        const T *sm = a.smorph + d.morph_off;
        for (int p = tid; p < n; p += nt) {
            const int by = p / Bx, bx = p - by * Bx, y = d.oy + by, x = d.ox + bx;
            double gm = 0.0;
            if ((unsigned)y < (unsigned)a.Ny && (unsigned)x < (unsigned)a.Nx) {
#pragma unroll
                for (int c = 0; c < SB_MAXC; ++c) {
                    if (c < C) {
                        const double g = grad_at<T>(a, s, c, y, x);
                        gm += sedv[c] * g;
                        gs[c] += g * (double)sm[p];
                    }
                }
            }
            zn[p] = (T)gm;
        }
        __syncthreads();
        const T *RCy = toep_vec<T>(a, d, 0), *ICy = toep_vec<T>(a, d, 1), *dRCy = toep_vec<T>(a, d, 2), *dICy = toep_vec<T>(a, d, 3);
        const T *RTx = toep_vec<T>(a, d, 4), *ITx = toep_vec<T>(a, d, 5), *dRTx = toep_vec<T>(a, d, 6), *dITx = toep_vec<T>(a, d, 7);
        toeplitz_apply<T, true>(zn, t1, t0, RCy, RTx, d.By, d.Bx, false, T(1));  // d/d image
        toeplitz_apply<T, true>(zn, t1, t0, ICy, ITx, d.By, d.Bx, true, T(-1));
        toeplitz_apply<T, true>(zn, t1, t2, dRCy, RTx, d.By, d.Bx, false, T(1)); // d/d s0 = <image, dCy^T g Tx>
        toeplitz_apply<T, true>(zn, t1, t2, dICy, ITx, d.By, d.Bx, true, T(-1));
        for (int p = tid; p < n; p += nt) gshift0 += (double)t2[p] * (double)mp[p];
        __syncthreads();
        toeplitz_apply<T, true>(zn, t1, t2, RCy, dRTx, d.By, d.Bx, false, T(1)); // d/d s1
        toeplitz_apply<T, true>(zn, t1, t2, ICy, dITx, d.By, d.Bx, true, T(-1));
        for (int p = tid; p < n; p += nt) gshift1 += (double)t2[p] * (double)mp[p];
        __syncthreads();
        gshift0 = block_sum(gshift0, red);
        gshift1 = block_sum(gshift1, red);
    }
    for (int p = tid; p < n; p += nt) {
        const int by = p / Bx, bx = p - by * Bx, y = d.oy + by, x = d.ox + bx;
        const T mval = mp[p];
        double gm = 0.0;
        if (d.shifting) {
            gm = (double)t0[p];
        } else if ((unsigned)y < (unsigned)a.Ny && (unsigned)x < (unsigned)a.Nx) {
#pragma unroll
            for (int c = 0; c < SB_MAXC; ++c) {
                if (c < C) {
                    const double g = grad_at<T>(a, s, c, y, x);
                    gm += sedv[c] * g;
                    gs[c] += g * (double)mval;
                }
            }
        }
        if (upd) {
            double m_ = (double)mm[p], v_ = (double)mv[p], vh_ = (double)mvh[p];
            const double psi = amsgrad(gm, m_, v_, vh_, it, a.fs);
            mm[p] = (T)m_, mv[p] = (T)v_, mvh[p] = (T)vh_;
            const T xn = (T)((double)mval - alpha * m_ / psi);
            gx[p] = xn;
            mp[p] = xn; // z0 = x
            zn[p] = xn; // first proximal argument: z0 - psi/max(psi) (z0 - x) = x exactly
            gp[p] = (T)psi;
            pmax = fmax(pmax, psi);
        }
        if (a.mode == 1 && a.g_morph) a.g_morph[d.morph_off + p] = gm;
    }
    // spectrum gradient: block reduction per band (double)
#pragma unroll
    for (int c = 0; c < SB_MAXC; ++c) {
        if (c < C) {
            const double t = block_sum(gs[c], red);
            if (tid == 0) gsum[c] = t;
        }
    }
    __syncthreads();
    if (a.mode == 1) {
        if (tid == 0 && a.g_sed)
            for (int c = 0; c < C; ++c) a.g_sed[(size_t)k * C + c] = gsum[c];
        if (tid == 0 && d.shifting && a.g_center) a.g_center[2 * d.point_idx] = gshift0, a.g_center[2 * d.point_idx + 1] = gshift1;
        return;
    }
    if (tid == 0 && d.shifting && a.mode == 0) { // the shift itself: AMSGrad step, no constraint (morphology.py:672-675)
        const double g2[2] = {gshift0, gshift1};
        double *cen = a.center + 2 * d.point_idx, *m = a.cen_m + 2 * d.point_idx, *v = a.cen_v + 2 * d.point_idx,
               *vh = a.cen_vhat + 2 * d.point_idx;
        for (int i = 0; i < 2; ++i) {
            double mm_ = m[i], vv = v[i], vvh = vh[i];
            const double psi = amsgrad(g2[i], mm_, vv, vvh, it, a.fs);
            m[i] = mm_, v[i] = vv, vh[i] = vvh;
            cen[i] = cen[i] - d.shift_step * mm_ / psi;
            if (!isfinite(cen[i])) atomicExch(a.status + s, SB_ERR_NONFINITE);
        }
    }
    if (upd) {
        const double psimax = block_max(pmax, red);
        bool bad = false;
        if (d.chain >= 0) {
            const DevChain &ch = a.chains[d.chain];
            const double gamma = alpha / psimax;
            const double fac = gamma / alpha;
            for (int sub = 0; sub < prox_max_iter_of(a, d.scene); ++sub) {
                if (sub > 0) {
                    for (int p = tid; p < n; p += nt) {
                        const double zz = (double)mp[p];
                        zn[p] = (T)(zz - fac * (double)gp[p] * (zz - (double)gx[p]));
                    }
                }
                __syncthreads();
                apply_chain<T>(zn, d.By, d.Bx, ch, a.monos, red);
                double dd = 0.0, nn = 0.0;
                bad = false;
                for (int p = tid; p < n; p += nt) {
                    const double zo = (double)mp[p], zv = (double)zn[p];
                    dd += (zv - zo) * (zv - zo);
                    nn += zo * zo;
                    mp[p] = zn[p];
                    bad |= !isfinite(zv);
                }
                dd = block_sum(dd, red);
                nn = block_sum(nn, red);
                if (dd <= a.fs.e_rel * a.fs.e_rel * nn) break;
            }
        } else {
            for (int p = tid; p < n; p += nt) bad |= !isfinite((double)mp[p]);
        }
        if (bad) atomicExch(a.status + s, SB_ERR_NONFINITE);
    }
    if (tid == 0 && !d.sed_fixed) sed_update<T>(a, d, k, gsum, it);
}

template <typename T> __global__ void __launch_bounds__(128) k_update(const UpdateArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int k = a.work ? a.work[blockIdx.x] : blockIdx.x;
    const DevSource &d = a.src[k];
    if (a.done[d.scene]) return;
    if (d.kind == 0) {
        update_extended<T>(a, d, k, smem);
    } else {
        double *red = reinterpret_cast<double *>(smem);
        update_point<T>(a, d, k, red);
    }
}


// ======================================================================================================
// K4-fast  grouped source update: the throughput path for image morphologies.
//
// A CTA holds G groups of 64 threads (2 warps); each group owns one source.  All groups of a CTA use the same
// constraint chain, hence the same radial-monotonicity operator, whose wavefront table is staged ONCE per CTA in
// shared memory (struct-of-arrays: neighbour indices, weights, pixel).  Only the image being projected lives in
// shared memory (the sweep and the symmetry partner need random access); the gradient-step result x, the metric
// psi and the running iterate z are streamed through L2 (coalesced, touched once per proximal sub-iteration).
// Groups synchronise with named barriers (bar.sync id, 64), so the 59-level sweep of one source never stalls the
// other seven.  Arithmetic is identical to update_extended (same formulas, same evaluation order).
// ======================================================================================================
// GT = threads per group (64 or 128), a template parameter of everything below

// Thread index inside a group.  Wavefront levels mostly hold fewer than 32 tasks, so only the warp with lt < 32 works during
// the sweep; a warp's scheduler is (warp index mod 4), and with the plain numbering those leading warps would all sit on
// schedulers 0 and 2 (GT = 64).  Every other pair of groups therefore numbers its warps the other way round.
// GT = 32: one warp per source -- no named barrier at all (__syncwarp orders the shared-memory traffic of the lanes), no idle
// second warp during the sweep, reductions by shuffles only.
template <int GT> __device__ __forceinline__ int group_lane(int g) {
    return GT == 64 ? (int)((threadIdx.x ^ (((unsigned)g >> 1 & 1u) << 5)) & 63u) : (int)(threadIdx.x & (GT - 1));
}
template <int GT> __device__ __forceinline__ void group_bar(int g) {
    if constexpr (GT == 32)
        __syncwarp();
    else
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GT) : "memory");
}
__device__ __forceinline__ double warp_sum_all(double v) { // every lane gets the sum (fixed butterfly order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct GroupRed {
    double *slot; // [2 parities][2 values][GT/32 warps]
    int g, par;
};
template <int GT> __device__ __forceinline__ void group_sum2(GroupRed &r, double &a, double &b) {
    constexpr int NW = GT / 32;
    if constexpr (GT == 32) {
        a = warp_sum_all(a);
        b = warp_sum_all(b);
        __syncwarp();
        return;
    }
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & (NW - 1);
    double *s = r.slot + r.par * 2 * NW;
    if (lane == 0) s[w] = a, s[NW + w] = b;
    group_bar<GT>(r.g);
    a = s[0], b = s[NW];
#pragma unroll
    for (int i = 1; i < NW; ++i) a += s[i], b += s[NW + i];
    r.par ^= 1; // the next reduction uses the other slot set; this one is rewritten only after another barrier
}
template <int GT> __device__ __forceinline__ double group_max(GroupRed &r, double a) {
    constexpr int NW = GT / 32;
    if constexpr (GT == 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        __syncwarp();
        return a;
    }
    a = warp_max(a);
    const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & (NW - 1);
    double *s = r.slot + r.par * 2 * NW;
    if (lane == 0) s[w] = a;
    group_bar<GT>(r.g);
    a = s[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) a = fmax(a, s[i]);
    r.par ^= 1;
    return a;
}

template <typename T> struct FastTable { // shared-memory image of a DevMono with nb == 4
    const uint2 *nbr;          // four neighbours per task as 16-bit BYTE offsets into the image (index * sizeof(T))
    const W4<T> *w;
    const unsigned short *pix; // byte offset of the task's own pixel
    const int *ls;
    int n_levels;
};

// 32-bit shared-window addresses + ld/st.shared: the sweep touches shared memory only through these, which spares the
// generic-to-shared address arithmetic the compiler otherwise redoes on every level (and one add per neighbour).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_i32(unsigned a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u32x2(unsigned a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_real(unsigned a, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_real(unsigned a, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_real(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void sts_real(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ W4<float> lds_w4(unsigned a, float) {
    W4<float> w;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.a), "=f"(w.b), "=f"(w.c), "=f"(w.d) : "r"(a));
    return w;
}
__device__ __forceinline__ W4<double> lds_w4(unsigned a, double) {
    W4<double> w;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w.a), "=d"(w.b) : "r"(a));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w.c), "=d"(w.d) : "r"(a + 16u));
    return w;
}

// Empty neighbour slots of the shared-memory table point at a spare cell behind the image that always holds 0 and carry
// weight 0, so the four products need no predicates: (+0) is added at the END of the reference's summation order
// (empty slots are trailing), which leaves the sum bit-identical.
template <typename T, int GT> __device__ __forceinline__ void group_sweep(T *img, const FastTable<T> &t, T min_gradient, int g) {
    const T keep = T(1) - min_gradient;
    const int lt = group_lane<GT>(g);
    const unsigned zb = smem_u32(img), a_nbr = smem_u32(t.nbr), a_w = smem_u32(t.w), a_pix = smem_u32(t.pix), a_ls = smem_u32(t.ls);
    int beg = lds_i32(a_ls);
    for (int L = 0; L < t.n_levels; ++L) {
        const int end = lds_i32(a_ls + 4u * (unsigned)(L + 1));
        for (int j = beg + lt; j < end; j += GT) {
            const uint2 nb = lds_u32x2(a_nbr + 8u * (unsigned)j);
            const W4<T> w = lds_w4(a_w + (unsigned)sizeof(W4<T>) * (unsigned)j, T(0));
            const unsigned ap = zb + lds_u16(a_pix + 2u * (unsigned)j);
            T ref = mul_rn(lds_real(zb + (nb.x & 0xffffu), T(0)), w.a);
            ref = add_rn(ref, mul_rn(lds_real(zb + (nb.x >> 16), T(0)), w.b));
            ref = add_rn(ref, mul_rn(lds_real(zb + (nb.y & 0xffffu), T(0)), w.c));
            ref = add_rn(ref, mul_rn(lds_real(zb + (nb.y >> 16), T(0)), w.d));
            const T cap = mul_rn(ref, keep);
            if (cap < lds_real(ap, T(0))) sts_real(ap, cap);
        }
        beg = end;
        group_bar<GT>(g);
    }
}

template <typename T, int GT>
__device__ void group_chain(T *a, int By, int Bx, const DevChain &ch, const FastTable<T> &tab, GroupRed &red) {
    const int n = By * Bx, g = red.g, lt = group_lane<GT>(g);
    for (int r = 0; r < ch.repeat; ++r) {
        for (int o = 0; o < ch.n_ops; ++o) {
            const sb_op op = ch.ops[o];
            switch (op.code) {
            case SB_OP_MONOTONIC:
                group_sweep<T, GT>(a, tab, (T)op.farg, g);
                break;
            case SB_OP_SYMMETRY: {
                const T hs = (T)(0.5 * op.farg), om = (T)(1.0 - op.farg);
                const int Hy = By + ((By & 1) == 0), Wx = Bx + ((Bx & 1) == 0);
                for (int p = lt; p < n; p += GT) {
                    const int y = p / Bx, x = p - y * Bx;
                    const int yr = Hy - 1 - y, xr = Wx - 1 - x;
                    if (yr >= By || xr >= Bx) {
                        const T u = a[p];
                        a[p] = hs * (u + T(0)) + om * u;
                    } else {
                        const int q = yr * Bx + xr;
                        if (q > p) {
                            const T u = a[p], v = a[q];
                            a[p] = hs * (u + v) + om * u;
                            a[q] = hs * (v + u) + om * v;
                        } else if (q == p) {
                            const T u = a[p];
                            a[p] = hs * (u + u) + om * u;
                        }
                    }
                }
                group_bar<GT>(g);
                break;
            }
            case SB_OP_POSITIVITY: {
                const T zero = (T)op.farg;
                for (int p = lt; p < n; p += GT) a[p] = a[p] < zero ? zero : a[p];
                group_bar<GT>(g);
                break;
            }
            case SB_OP_CENTER_ON: {
                if (lt == 0) {
                    const int c = (By / 2) * Bx + Bx / 2;
                    const T tiny = (T)op.farg;
                    a[c] = a[c] < tiny ? tiny : a[c];
                }
                group_bar<GT>(g);
                break;
            }
            case SB_OP_NORMALIZE: {
                double acc, dummy = 0.0;
                if (op.iarg == 1) {
                    acc = -INFINITY;
                    for (int p = lt; p < n; p += GT) acc = fmax(acc, (double)a[p]);
                    acc = group_max<GT>(red, acc);
                } else {
                    acc = 0.0;
                    for (int p = lt; p < n; p += GT) acc += (double)a[p];
                    group_sum2<GT>(red, acc, dummy);
                }
                const T den = (T)acc;
                for (int p = lt; p < n; p += GT) a[p] = a[p] / den;
                group_bar<GT>(g);
                break;
            }
            default:
                break;
            }
        }
    }
}

#define SB_FAST_MAXC 8

// group-wide maximum in the image type (exact: a maximum needs no extra precision)
template <typename T, int GT> __device__ __forceinline__ T group_max_t(GroupRed &r, T a) {
    constexpr int NW = GT / 32;
    if constexpr (GT == 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T b = __shfl_xor_sync(0xffffffffu, a, o);
            a = b > a ? b : a;
        }
        __syncwarp();
        return a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T b = __shfl_down_sync(0xffffffffu, a, o);
        a = b > a ? b : a;
    }
    const int lane = threadIdx.x & 31, w = (threadIdx.x >> 5) & (NW - 1);
    T *s = reinterpret_cast<T *>(r.slot + r.par * 2 * NW);
    if (lane == 0) s[w] = a;
    group_bar<GT>(r.g);
    a = s[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) a = s[i] > a ? s[i] : a;
    r.par ^= 1;
    return a;
}

// The ExtendedSource chain Monotonicity -> [Symmetry] -> Positivity -> CenterOn -> Normalization("max")
// (morphology.py:644-669) on an odd x odd box, recognised so that everything after the sweep runs in two passes.
struct FusedChain {
    int ok, has_sym;
    double mono_grad, sym, zero, tiny;
};
__device__ inline FusedChain fused_chain_of(const DevChain &ch, int By, int Bx) {
    FusedChain f;
    f.ok = 0, f.has_sym = 0, f.mono_grad = 0, f.sym = 0, f.zero = 0, f.tiny = 0;
    if (ch.repeat != 1 || !(By & 1) || !(Bx & 1)) return f;
    int i = 0;
    if (i >= ch.n_ops || ch.ops[i].code != SB_OP_MONOTONIC) return f;
    f.mono_grad = ch.ops[i++].farg;
    if (i < ch.n_ops && ch.ops[i].code == SB_OP_SYMMETRY) f.has_sym = 1, f.sym = ch.ops[i++].farg;
    if (i >= ch.n_ops || ch.ops[i].code != SB_OP_POSITIVITY) return f;
    f.zero = ch.ops[i++].farg;
    if (i >= ch.n_ops || ch.ops[i].code != SB_OP_CENTER_ON) return f;
    f.tiny = ch.ops[i++].farg;
    if (i >= ch.n_ops || ch.ops[i].code != SB_OP_NORMALIZE || ch.ops[i].iarg != 1) return f;
    f.ok = (i + 1 == ch.n_ops);
    return f;
}

// MAXT: upper bound of the CTA size the launch will use (threads = GT x groups).  The kernel always runs one CTA per SM, so a
// smaller bound hands the spare registers of the file to each thread (832 threads: 78 registers instead of 64).
template <typename T, int GT, int MAXT = 1024> __global__ void __launch_bounds__(MAXT, 1) k_update_fast(const UpdateArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int G = a.fast_G, g = threadIdx.x / GT, lt = group_lane<GT>(g);
    const int *mine = a.fast_groups + (size_t)blockIdx.x * G;
    // ---- shared memory carve-up: table (W4 | uint2 | int ls | u16 pix), reduction slots, G images
    const int cap = a.fast_table_cap;
    W4<T> *s_w = reinterpret_cast<W4<T> *>(smem);
    uint2 *s_nbr = reinterpret_cast<uint2 *>(s_w + cap);
    double *s_red = reinterpret_cast<double *>(s_nbr + cap);
    double *s_gsum = s_red + 4 * (GT / 32) * G; // [G][SB_MAXC]
    int *s_ls = reinterpret_cast<int *>(s_gsum + SB_MAXC * G); // [512]
    unsigned short *s_pix = reinterpret_cast<unsigned short *>(s_ls + 512);
    T *s_img = reinterpret_cast<T *>(s_pix + ((cap + 7) & ~7));

    // ---- the CTA's common operator table -> shared memory (first live group defines the chain)
    int k0 = -1;
    for (int i = 0; i < G; ++i)
        if (mine[i] >= 0) {
            k0 = mine[i];
            break;
        }
    FastTable<T> tab;
    tab.nbr = s_nbr, tab.w = s_w, tab.pix = s_pix, tab.ls = s_ls, tab.n_levels = 0;
    const DevChain &ch = a.chains[a.src[k0].chain];
    for (int o = 0; o < ch.n_ops; ++o)
        if (ch.ops[o].code == SB_OP_MONOTONIC) {
            const DevMono &mo = a.monos[ch.ops[o].iarg];
            const uint2 *gn = reinterpret_cast<const uint2 *>(mo.code);
            const W4<T> *gw = reinterpret_cast<const W4<T> *>(mo.w);
            const unsigned spare = (unsigned)mo.n_pix; // index of the always-zero cell behind each image
            for (int j = threadIdx.x; j < mo.n_tasks; j += blockDim.x) {
                const uint2 v = gn[j];
                unsigned i0 = v.x & 0xffffu, i1 = v.x >> 16, i2 = v.y & 0xffffu, i3 = v.y >> 16;
                i0 = (i0 == 0xffffu ? spare : i0) * (unsigned)sizeof(T), i1 = (i1 == 0xffffu ? spare : i1) * (unsigned)sizeof(T);
                i2 = (i2 == 0xffffu ? spare : i2) * (unsigned)sizeof(T), i3 = (i3 == 0xffffu ? spare : i3) * (unsigned)sizeof(T);
                s_nbr[j] = make_uint2(i0 | (i1 << 16), i2 | (i3 << 16)); // byte offsets (host: (n_pix + 1) sizeof(T) < 65536)
                s_w[j] = gw[j];
                s_pix[j] = (unsigned short)(mo.pix[j] * (int)sizeof(T));
            }
            for (int j = threadIdx.x; j <= mo.n_levels; j += blockDim.x) s_ls[j] = mo.level_start[j];
            tab.n_levels = mo.n_levels;
        }
    __syncthreads();
    if (g >= G) return;
    const int k = mine[g];
    if (k < 0) return;
    const DevSource &d = a.src[k];
    const int s = d.scene;
    if (a.done[s]) return;

    const int it = a.it_ptr[s], C = a.C, n = d.By * d.Bx, Bx = d.Bx;
    const unsigned magic = 0xffffffffu / (unsigned)Bx + 1u; // p / Bx == umulhi(p, magic) for p, Bx < 65536
    T *zn = s_img + (size_t)g * a.fast_npix;
    GroupRed red;
    red.slot = s_red + 4 * (GT / 32) * g, red.g = g, red.par = 0;
    double *gsum = s_gsum + SB_MAXC * g; // first holds the spectrum (read-only), then the spectrum gradient

    T gs[SB_FAST_MAXC]; // per-thread partial sums in T (<= 27 terms each), reduced in double
#pragma unroll
    for (int c = 0; c < SB_FAST_MAXC; ++c) gs[c] = T(0);
    if (lt < C) gsum[lt] = a.sed[(size_t)k * C + lt];
    if (lt == 0) zn[n] = T(0); // the spare cell of group_sweep
    group_bar<GT>(g);
    T *mp = a.morph + d.morph_off, *mm = a.morph_m + d.morph_off, *mv = a.morph_v + d.morph_off,
      *mvh = a.morph_vhat + d.morph_off, *xs = a.scratch_x + d.morph_off, *ps = a.scratch_ps + d.morph_off;
    const double alpha = d.morph_step;
    const bool upd = !d.morph_fixed;
    double pmax = 0.0;
    constexpr int PB = 2; // pixels per trip: all loads first (see pass B below); 3 is slower even with 78 registers
    for (int p0 = lt; p0 < n; p0 += PB * GT) {
        T mval[PB], m0[PB], v0[PB], vh0[PB];
        double gm[PB];
        if (a.n_obs == 1) { // every load of the trip first, then the arithmetic
            T gv[PB][SB_FAST_MAXC];
            bool in[PB];
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                const int p = p0 + i * GT;
                mval[i] = m0[i] = v0[i] = vh0[i] = T(0);
                in[i] = false;
                if (p < n) {
                    const int by = (int)__umulhi((unsigned)p, magic), bx = p - by * Bx, y = d.oy + by, x = d.ox + bx;
                    mval[i] = mp[p];
                    if (upd) m0[i] = mm[p], v0[i] = mv[p], vh0[i] = mvh[p];
                    in[i] = (unsigned)y < (unsigned)a.Ny && (unsigned)x < (unsigned)a.Nx;
                    if (in[i]) grad_bands<T, SB_FAST_MAXC>(a, s, C, y, x, gv[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                gm[i] = 0.0;
                if (in[i]) {
#pragma unroll
                    for (int c = 0; c < SB_FAST_MAXC; ++c) {
                        if (c < C) {
                            const double gg = 0.0 + (double)gv[i][c];
                            gm[i] += gsum[c] * gg;
                            gs[c] += (T)gg * mval[i];
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                const int p = p0 + i * GT;
                mval[i] = m0[i] = v0[i] = vh0[i] = T(0);
                gm[i] = 0.0;
                if (p < n) {
                    const int by = (int)__umulhi((unsigned)p, magic), bx = p - by * Bx, y = d.oy + by, x = d.ox + bx;
                    mval[i] = mp[p];
                    if (upd) m0[i] = mm[p], v0[i] = mv[p], vh0[i] = mvh[p];
                    if ((unsigned)y < (unsigned)a.Ny && (unsigned)x < (unsigned)a.Nx) {
#pragma unroll
                        for (int c = 0; c < SB_FAST_MAXC; ++c) {
                            if (c < C) {
                                const double gg = grad_at<T>(a, s, c, y, x);
                                gm[i] += gsum[c] * gg;
                                gs[c] += (T)gg * mval[i];
                            }
                        }
                    }
                }
            }
        }
        if (upd) {
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                const int p = p0 + i * GT;
                if (p < n) {
                    double m_ = (double)m0[i], v_ = (double)v0[i], vh_ = (double)vh0[i];
                    const double psi = amsgrad(gm[i], m_, v_, vh_, it, a.fs);
                    mm[p] = (T)m_, mv[p] = (T)v_, mvh[p] = (T)vh_;
                    const T xn = (T)((double)mval[i] - alpha * m_ / psi);
                    xs[p] = xn;
                    mp[p] = xn; // z0 = x
                    zn[p] = xn; // first proximal argument: z0 - psi/max(psi) (z0 - x) = x exactly
                    ps[p] = (T)psi;
                    pmax = fmax(pmax, psi);
                }
            }
        }
    }
    // spectrum gradient (pairs of bands per reduction); every thread is past its reads of the spectrum in gsum
    group_bar<GT>(g);
#pragma unroll
    for (int c = 0; c < SB_FAST_MAXC; c += 2) {
        if (c < C) {
            double u = (double)gs[c], v = (double)gs[c + 1];
            group_sum2<GT>(red, u, v);
            if (lt == 0) {
                gsum[c] = u;
                if (c + 1 < C) gsum[c + 1] = v;
            }
        }
    }
    if (upd) {
        const double psimax = group_max<GT>(red, pmax);
        bool bad = false;
        if (d.chain >= 0) {
            const double gamma = alpha / psimax;
            const double fac = gamma / alpha;
            const FusedChain fc = fused_chain_of(ch, d.By, d.Bx);
            const double e2 = a.fs.e_rel * a.fs.e_rel;
            if (fc.ok) {
                const T hs = (T)(0.5 * fc.sym), om = (T)(1.0 - fc.sym), zero = (T)fc.zero, tiny = (T)fc.tiny;
                const int half = (n - 1) >> 1; // centre pixel index (odd x odd box): its 180-degree partner is itself
                for (int sub = 0; sub < prox_max_iter_of(a, d.scene); ++sub) {
                    group_sweep<T, GT>(zn, tab, (T)fc.mono_grad, g);
                    // pass A: symmetry (pairs p, n-1-p), positivity, centre floor, running maximum
                    T mx = -INFINITY;
                    for (int p0 = lt; p0 <= half; p0 += 2 * GT) { // two pairs per trip, their four loads first
                        T uu[2], vv[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int p = p0 + i * GT;
                            uu[i] = vv[i] = T(0);
                            if (p <= half) uu[i] = zn[p], vv[i] = zn[n - 1 - p];
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int p = p0 + i * GT;
                            if (p <= half) {
                                T u = uu[i], v = vv[i];
                                if (fc.has_sym) {
                                    const T un = hs * (u + v) + om * u, vn = hs * (v + u) + om * v;
                                    u = un, v = vn;
                                }
                                u = u < zero ? zero : u; // np.maximum / max(): a NaN survives and is caught below
                                v = v < zero ? zero : v;
                                if (p == half) {
                                    u = u < tiny ? tiny : u;
                                    v = u;
                                }
                                zn[p] = u, zn[n - 1 - p] = v;
                                mx = u > mx ? u : mx;
                                mx = v > mx ? v : mx;
                            }
                        }
                    }
                    const T den = group_max_t<T, GT>(red, mx); // barrier inside: pass A is complete for the whole group
                    // pass B: normalise, convergence sums, store z, next proximal argument
                    double dd = 0.0, nn = 0.0;
                    bad = false;
                    const bool last = sub + 1 == prox_max_iter_of(a, d.scene);
                    // (loads of a batch are issued before its stores: the compiler cannot reorder them itself because
                    // mp, ps and xs may alias as far as it knows, and one L2 round trip per pixel would dominate)
                    for (int p0 = lt; p0 < n; p0 += 4 * GT) {
                        T zr[4], zo_[4], ps_[4], xs_[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int p = p0 + i * GT;
                            zr[i] = zo_[i] = ps_[i] = xs_[i] = T(0);
                            if (p < n) {
                                zr[i] = zn[p];
                                zo_[i] = mp[p];
                                if (!last) ps_[i] = ps[p], xs_[i] = xs[p];
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int p = p0 + i * GT;
                            if (p < n) {
                                const T r = zr[i] / den;
                                const double zo = (double)zo_[i], zv = (double)r;
                                dd += (zv - zo) * (zv - zo);
                                nn += zo * zo;
                                mp[p] = r;
                                bad |= !isfinite(zv);
                                if (!last) zn[p] = (T)(zv - fac * (double)ps_[i] * (zv - (double)xs_[i]));
                            }
                        }
                    }
                    group_sum2<GT>(red, dd, nn); // barrier inside: zn is complete before the next sweep
                    if (a.prox_hist && lt == 0 && (dd <= e2 * nn || last)) atomicAdd(a.prox_hist + min(sub + 1, 15), 1ull);
                    if (dd <= e2 * nn) break;
                }
            } else {
                for (int sub = 0; sub < prox_max_iter_of(a, d.scene); ++sub) {
                    if (sub > 0) {
                        for (int p = lt; p < n; p += GT) {
                            const double zz = (double)mp[p];
                            zn[p] = (T)(zz - fac * (double)ps[p] * (zz - (double)xs[p]));
                        }
                    }
                    group_bar<GT>(g);
                    group_chain<T, GT>(zn, d.By, d.Bx, ch, tab, red);
                    double dd = 0.0, nn = 0.0;
                    bad = false;
                    for (int p = lt; p < n; p += GT) {
                        const double zo = (double)mp[p], zv = (double)zn[p];
                        dd += (zv - zo) * (zv - zo);
                        nn += zo * zo;
                        mp[p] = zn[p];
                        bad |= !isfinite(zv);
                    }
                    group_sum2<GT>(red, dd, nn);
                    if (dd <= e2 * nn) break;
                }
            }
        } else {
            for (int p = lt; p < n; p += GT) bad |= !isfinite((double)mp[p]);
        }
        if (bad) atomicExch(a.status + s, SB_ERR_NONFINITE);
    }
    group_bar<GT>(g); // gsum visible to the updating thread
    if (lt == 0 && !d.sed_fixed) sed_update<T>(a, d, k, gsum, it);
}

// ======================================================================================================
// K7  per-scene loss reduction + stop rule + iteration counter      reference: blend.py:264-274, 276-302
// ======================================================================================================
// per-scene run state (device array, read by the host between slices of a fit)
enum { SB_RUN = 0, SB_CONVERGED = 1, SB_PAUSED = 2, SB_EXHAUSTED = 4, SB_FAILED = 8, SB_CONV_PENDING = 16 }; // distinct bits

struct LossArgs {
    int n_obs;
    const double *partials[SB_MAX_OBS]; // per observation: [S][n_part[o]]
    int n_part[SB_MAX_OBS];
    const double *loss_const; // [S]
    double *loss;             // [S][cap]
    int cap;
    int *done, *n_iter, *n_active_next;
    int *it_arr;              // [S] iteration counter of the running adaprox call; advanced here (last kernel of an iteration)
    const int *limit;         // [S] optional: stop when the loss history reaches this length (max_iter of the scene's fit)
    int *state;               // [S] SB_RUN ... (| SB_CONV_PENDING)
    const int *status;
    int advance;              // 0: evaluate only (counters untouched)
    FitScalars fs;
};

// one block per scene: fixed-order reduction of the chi^2 partials -> loss[s][len], then Blend._callback's decisions
// (blend.py:276-302): non-finite parameters, inspection of the sources every ``pause_every`` iterations (the host runs
// src.update() on a paused scene and either restarts or resumes it), the stop rule, the iteration budget.
__global__ void __launch_bounds__(128) k_loss_stop(const LossArgs a) {
    __shared__ double red[40];
    const int s = blockIdx.x;
    if (a.done[s]) return;
    double acc = 0.0;
    for (int o = 0; o < a.n_obs; ++o) {
        const double *p = a.partials[o] + (size_t)s * a.n_part[o];
        for (int i = threadIdx.x; i < a.n_part[o]; i += blockDim.x) acc += p[i];
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        const int it = a.it_arr[s], idx = a.n_iter[s]; // idx: length of this fit's loss history so far
        const double l1 = a.loss_const[s] + 0.5 * acc;
        a.loss[(size_t)s * a.cap + idx] = l1;
        if (!a.advance) {
            a.n_iter[s] = idx + 1;
            return;
        }
        a.n_iter[s] = idx + 1;
        a.it_arr[s] = it + 1;
        bool conv = false;
        if (!a.fs.fixed_iterations && it > 0 && it > a.fs.min_iter && idx > 0) {
            const double l0 = a.loss[(size_t)s * a.cap + idx - 1];
            conv = fabs(l1 - l0) < a.fs.e_rel * fabs(l1);
        }
        int st = SB_RUN;
        if (a.status[s] != 0)
            st = SB_FAILED;
        else if (a.fs.pause_every > 0 && it > 0 && it % a.fs.pause_every == 0)
            st = SB_PAUSED | (conv ? SB_CONV_PENDING : 0);
        else if (conv)
            st = SB_CONVERGED;
        else if (a.limit && idx + 1 >= a.limit[s])
            st = SB_EXHAUSTED;
        a.state[s] = st;
        if (st != SB_RUN)
            a.done[s] = 1;
        else
            atomicAdd(a.n_active_next, 1);
    }
}

__global__ void k_tick(int *n_active, int *n_active_next) {
    *n_active = *n_active_next;
    *n_active_next = 0;
}

// ======================================================================================================
// Dynamic boxes: ImageMorphology.update's decision for one source (morphology.py:52-68 shrink_box, 132-207 update;
// initialization.py:173-177 allowed sizes 21, 31, 41, ...).  One CTA per source; nothing is modified.
//   shrink: d = number of complete outer rings with every pixel <= 0; new size = smallest allowed >= size - 2 d, if smaller
//   grow  : gu = -m / sqrt(sqrt(v)) * step over pixels with v != 0; pull = gu where image > 0 else 0; if the mean pull of one
//           of the four edges exceeds 0.1: new size = smallest allowed >= size + 1
// The host evaluates the same rules in float64 from the same values; its edge means are pairwise sums, so a mean within 1e-9
// (relative) of the threshold is reported as "undecided" (-1) and left to the host.
// ======================================================================================================
__device__ __forceinline__ int minimal_boxsize_dev(int size) {
    int b = 21;
    while (b < size) b += 10;
    return b;
}
template <typename T>
__global__ void __launch_bounds__(128) k_inspect(const DevSource *src, int n_src, const T *morph, const T *morph_m, const T *morph_v,
                                                 const int *state, int *action) {
    __shared__ int ring_ne[128];
    __shared__ double red[40];
    const int k = blockIdx.x;
    if (k >= n_src) return;
    const DevSource &d = src[k];
    const int tid = threadIdx.x, nt = blockDim.x;
    int act = 0;
    const bool look = d.kind == 0 && d.resizing && !d.morph_fixed && (state[d.scene] & SB_PAUSED);
    if (!look) {
        if (tid == 0) action[k] = 0;
        return;
    }
    if (d.By != d.Bx || d.By > 250) { // the rules are written for square boxes; anything else is the host's call
        if (tid == 0) action[k] = -1;
        return;
    }
    const int n = d.By, np = n * n;
    const T *x = morph + d.morph_off, *m = morph_m + d.morph_off, *v = morph_v + d.morph_off;
    for (int r = tid; r < 128; r += nt) ring_ne[r] = 0;
    __syncthreads();
    for (int p = tid; p < np; p += nt) {
        const int yy = p / n, xx = p - yy * n;
        const int r = min(min(yy, xx), min(n - 1 - yy, n - 1 - xx));
        if (!(x[p] <= T(0))) ring_ne[r] = 1; // np.all(image <= 0) fails for this ring (a NaN fails it too)
    }
    __syncthreads();
    int dist = 0;
    while (dist < (n + 1) / 2 && !ring_ne[dist]) ++dist;
    const int shrunk = minimal_boxsize_dev(n - 2 * dist);
    if (shrunk < n) {
        if (tid == 0) action[k] = shrunk;
        return;
    }
    // edge pull: four edges, masked means
    double sum[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0};
    for (int i = tid; i < n; i += nt) {
        const int idx[4] = {i * n, i * n + n - 1, i, (n - 1) * n + i}; // [:,0], [:,-1], [0,:], [-1,:]
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double vv = (double)v[idx[e]];
            if (vv != 0.0) {
                const double gu = -(double)m[idx[e]] / sqrt(sqrt(vv)) * d.morph_step;
                sum[e] += x[idx[e]] > T(0) ? gu : gu * 0.0;
                cnt[e] += 1.0;
            }
        }
    }
    bool grow = false, unsure = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double s_ = block_sum(sum[e], red), c_ = block_sum(cnt[e], red);
        if (c_ > 0) {
            const double mean = s_ / c_;
            if (fabs(mean - 0.1) < 1e-9) unsure = true;
            if (mean > 0.1) grow = true;
            if (!(mean == mean)) unsure = true; // NaN: let the host see it
        }
    }
    act = unsure ? -1 : (grow ? minimal_boxsize_dev(n + 1) : 0);
    if (tid == 0) action[k] = act;
}

// ======================================================================================================
// single-operator kernels (test / plugin entry points)
// ======================================================================================================
template <typename T>
__global__ void __launch_bounds__(128) k_chain_only(T *img, int By, int Bx, const DevChain *ch, const DevMono *monos) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = By * Bx, npad = (n + 1) & ~1;
    T *a = reinterpret_cast<T *>(smem);
    double *red = reinterpret_cast<double *>(a + npad + (sizeof(T) == 4 ? (npad & 2) : 0));
    T *g = img + (size_t)blockIdx.x * n;
    for (int p = threadIdx.x; p < n; p += blockDim.x) a[p] = g[p];
    __syncthreads();
    apply_chain<T>(a, By, Bx, *ch, monos, red);
    for (int p = threadIdx.x; p < n; p += blockDim.x) g[p] = a[p];
}

// [rows][w] complex128 -> [rows][pitch] complex T, scaled
template <typename T>
__global__ void k_cast_scale_cplx_pitched(const double2 *in, typename Cx<T>::type *out, long long rows, int w, int pitch, double scale) {
    const long long n = rows * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / w;
        const int k = (int)(i - r * w);
        typename Cx<T>::type v;
        v.x = (T)(in[i].x * scale);
        v.y = (T)(in[i].y * scale);
        out[r * pitch + k] = v;
    }
}
template <typename T> __global__ void k_cast_scale_cplx(const double2 *in, typename Cx<T>::type *out, long long n, double scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        typename Cx<T>::type r;
        r.x = (T)(in[i].x * scale);
        r.y = (T)(in[i].y * scale);
        out[i] = r;
    }
}
// kernel image [n_img][Py][Px] -> zeroed grid [n_img][Fy][Fx] with pixel (i,j) at ((i+y0) mod Fy, (j+x0) mod Fx):
// the reference's centre-pad + ifftshift placement (fft.py:82-113, 255-273)
__global__ void k_embed_kernel(const double *ker, double *grid, int n_img, int Py, int Px, int Fy, int Fx, int y0, int x0) {
    const long long total = (long long)n_img * Py * Px;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = i % Px;
        const long long t = i / Px;
        const int y = t % Py;
        const long long im = t / Py;
        int gy = (y + y0) % Fy, gx = (x + x0) % Fx;
        if (gy < 0) gy += Fy;
        if (gx < 0) gx += Fx;
        grid[(im * Fy + gy) * Fx + gx] = ker[i];
    }
}
// real-space filter of the reference's apply_filter (operators_pybind11.cc:39-56), gather form: one thread per output pixel
template <typename T>
__global__ void k_apply_filter(const T *img, int H, int W, const T *values, const int *ys, const int *ye, const int *xs, const int *xe,
                               int n_taps, T *out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    T acc = T(0);
    for (int n = 0; n < n_taps; ++n) {
        const int r = y - ys[n], c = x - xs[n];
        if (r >= 0 && c >= 0 && r < H - ys[n] - ye[n] && c < W - xs[n] - xe[n])
            acc = add_rn(acc, mul_rn(values[n], img[(size_t)(r + ye[n]) * W + c + xe[n]]));
    }
    out[(size_t)y * W + x] = acc;
}
template <typename TI, typename TO> __global__ void k_cast(const TI *in, TO *out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (TO)in[i];
}
// copy the [0:Ny,0:Nx) corner of padded grids to a dense cube and back
template <typename T> __global__ void k_crop(const T *grid, T *out, int n_img, int Fy, int Fx, int Ny, int Nx) {
    const long long total = (long long)n_img * Ny * Nx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = i % Nx;
        const long long t = i / Nx;
        const int y = t % Ny;
        const long long im = t / Ny;
        out[i] = grid[(im * Fy + y) * Fx + x];
    }
}
template <typename T> __global__ void k_embed(const T *in, T *grid, int n_img, int Fy, int Fx, int Ny, int Nx) {
    const long long total = (long long)n_img * Ny * Nx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = i % Nx;
        const long long t = i / Nx;
        const int y = t % Ny;
        const long long im = t / Ny;
        grid[(im * Fy + y) * Fx + x] = in[i];
    }
}

} // namespace sb
