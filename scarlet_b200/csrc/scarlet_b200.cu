// scarlet_b200 -- host side of the CUDA library: plan (device memory, cuFFT plans, CUDA graph of one
// proximal-gradient iteration) and the C ABI declared in include/scarlet_b200.h.
//
// The plan replaces the closures that scarlet's Blend.fit hands to proxmin.adaprox
// (reference scarlet/blend.py:103-180): gradient of the loss, step sizes, proximal operators, callback.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>

#include "kernels.cuh"
#include "spectral.cuh"
#include "update_warp.cuh"
#include "psf_shift.cuh"

namespace sb {

thread_local std::string g_err;
int set_err(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

static const char *kStageNames[SB_N_STAGES] = {"render",       "fft_fwd_model", "kmul",     "fft_inv_model", "residual_loss",
                                               "fft_fwd_resid", "kmul_conj",     "fft_inv_grad", "source_update", "advance"};

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    // (contents are unspecified after alloc; an allocation that is large enough is kept -- cudaFree synchronises the device,
    // and a batch with dynamic boxes re-sizes its source-side buffers at every re-plan)
    int alloc(size_t count) {
        if (p && cap >= count && count > 0) {
            n = count;
            return SB_OK;
        }
        release();
        n = count;
        if (count == 0) return SB_OK;
        SB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
        cap = count;
        return SB_OK;
    }
    int zero(cudaStream_t st) {
        if (n) SB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
        return SB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0, cap = 0;
    }
    ~DevBuf() { release(); }
    size_t bytes() const { return (cap > n ? cap : n) * sizeof(T); }
};

// ---- wavefront tables ---------------------------------------------------------------------------
struct HostMono {
    int n_pix = 0, n_tasks = 0, n_levels = 0, nb = 4;
    int off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<int> pix, level_start;
    std::vector<uint16_t> nbr; // [nb][n_tasks] (grouped by 4 below when uploaded)
    std::vector<double> w;     // [nb][n_tasks]
};

// Turn the arguments of the reference's sequential sweep into a level schedule that is exact for ANY input
// (read-after-write, write-after-write and write-after-read hazards between tasks are all honoured).
template <typename T> static int build_mono(const sb_mono_desc &d, HostMono &h) {
    if (d.n_pix <= 0 || d.n_pix >= 0xffff) return set_err(SB_ERR_ARG, "monotonic operator: n_pix=%d out of range", d.n_pix);
    if (d.n_off <= 0 || d.n_off > 8) return set_err(SB_ERR_ARG, "monotonic operator: n_off=%d (max 8)", d.n_off);
    if (!d.weights || !d.offsets || (d.n_idx > 0 && !d.dist_idx)) return set_err(SB_ERR_ARG, "monotonic operator: null table");
    h.n_pix = d.n_pix;
    h.n_tasks = d.n_idx;
    for (int i = 0; i < d.n_off; ++i) h.off[i] = d.offsets[i];
    std::vector<int> last_w(d.n_pix, 0), last_r(d.n_pix, 0), level(d.n_idx, 0), cnt(d.n_idx, 0);
    std::vector<uint16_t> nb8((size_t)8 * d.n_idx, 0xffff);
    std::vector<double> w8((size_t)8 * d.n_idx, 0.0);
    int maxcnt = 0, maxlvl = 0;
    for (int t = 0; t < d.n_idx; ++t) {
        const int p = d.dist_idx[t];
        if (p < 0 || p >= d.n_pix) return set_err(SB_ERR_ARG, "monotonic operator: dist_idx[%d]=%d out of range", t, p);
        int lvl = std::max(last_w[p], last_r[p]);
        int c = 0;
        for (int i = 0; i < d.n_off; ++i) {
            const T wt = (T)d.weights[(size_t)i * d.n_pix + p]; // the reference tests the weight in the image dtype
            if (wt > 0) {
                const int q = p + d.offsets[i];
                if (q < 0 || q >= d.n_pix) return set_err(SB_ERR_ARG, "monotonic operator: neighbour of pixel %d outside the image", p);
                lvl = std::max(lvl, last_w[q]);
                nb8[(size_t)c * d.n_idx + t] = (uint16_t)q;
                w8[(size_t)c * d.n_idx + t] = (double)wt;
                ++c;
            }
        }
        lvl += 1;
        for (int i = 0; i < c; ++i) {
            const int q = nb8[(size_t)i * d.n_idx + t];
            last_r[q] = std::max(last_r[q], lvl);
        }
        last_w[p] = lvl;
        level[t] = lvl;
        cnt[t] = c;
        maxcnt = std::max(maxcnt, c);
        maxlvl = std::max(maxlvl, lvl);
    }
    h.nb = maxcnt <= 4 ? 4 : 8;
    h.n_levels = maxlvl;
    // stable counting sort by level
    h.level_start.assign(maxlvl + 1, 0);
    for (int t = 0; t < d.n_idx; ++t) h.level_start[level[t]]++; // level in 1..maxlvl -> slot level
    {
        int run = 0;
        for (int l = 1; l <= maxlvl; ++l) {
            const int c = h.level_start[l];
            h.level_start[l - 1] = run;
            run += c;
        }
        h.level_start[maxlvl] = run;
    }
    std::vector<int> cursor(h.level_start.begin(), h.level_start.end());
    h.pix.assign(d.n_idx, 0);
    h.nbr.assign((size_t)h.nb * d.n_idx, 0xffff);
    h.w.assign((size_t)h.nb * d.n_idx, 0.0);
    for (int t = 0; t < d.n_idx; ++t) {
        const int j = cursor[level[t] - 1]++;
        h.pix[j] = d.dist_idx[t];
        for (int i = 0; i < cnt[t]; ++i) {
            h.nbr[(size_t)i * d.n_idx + j] = nb8[(size_t)i * d.n_idx + t];
            h.w[(size_t)i * d.n_idx + j] = w8[(size_t)i * d.n_idx + t];
        }
    }
    return SB_OK;
}

template <typename T> struct MonoDevice {
    DevBuf<int> pix, level_start;
    DevBuf<uint2> nbr;
    DevBuf<W4<T>> w;
    DevBuf<unsigned char> wtab;
    DevMono dev;
    int upload(const HostMono &h, cudaStream_t st) {
        const int n = h.n_tasks, groups = h.nb / 4;
        SB_TRY(pix.alloc(std::max(n, 1)));
        SB_TRY(level_start.alloc(h.level_start.size() + 1));
        SB_TRY(nbr.alloc((size_t)groups * std::max(n, 1)));
        SB_TRY(w.alloc((size_t)groups * std::max(n, 1)));
        std::vector<uint2> hn((size_t)groups * std::max(n, 1));
        std::vector<W4<T>> hw((size_t)groups * std::max(n, 1));
        for (int g = 0; g < groups; ++g)
            for (int j = 0; j < n; ++j) {
                const uint16_t *q = &h.nbr[0];
                const size_t s0 = (size_t)(4 * g) * n + j, s1 = s0 + n, s2 = s1 + n, s3 = s2 + n;
                uint2 u;
                u.x = (unsigned)q[s0] | ((unsigned)q[s1] << 16);
                u.y = (unsigned)q[s2] | ((unsigned)q[s3] << 16);
                hn[(size_t)g * n + j] = u;
                W4<T> ww;
                ww.a = (T)h.w[s0], ww.b = (T)h.w[s1], ww.c = (T)h.w[s2], ww.d = (T)h.w[s3];
                hw[(size_t)g * n + j] = ww;
            }
        if (n) {
            SB_CUDA(cudaMemcpyAsync(pix.p, h.pix.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
            SB_CUDA(cudaMemcpyAsync(nbr.p, hn.data(), hn.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
            SB_CUDA(cudaMemcpyAsync(w.p, hw.data(), hw.size() * sizeof(W4<T>), cudaMemcpyHostToDevice, st));
        }
        SB_CUDA(cudaMemcpyAsync(level_start.p, h.level_start.data(), h.level_start.size() * sizeof(int),
                                cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaStreamSynchronize(st)); // host staging vectors die here
        dev.n_pix = h.n_pix, dev.n_tasks = n, dev.n_levels = h.n_levels, dev.nb = h.nb;
        for (int i = 0; i < 8; ++i) dev.off[i] = h.off[i];
        dev.pix = pix.p, dev.code = reinterpret_cast<const unsigned *>(nbr.p), dev.w = w.p, dev.level_start = level_start.p;
        dev.wtab = nullptr, dev.w_trips = 0, dev.w_cap = 0;
        if (h.nb == 4 && (size_t)(h.n_pix + 1) * sizeof(T) <= 65535) { // table by trips (see DevMono::wtab)
            int trips = 0;
            for (int L = 0; L < h.n_levels; ++L) trips += (h.level_start[L + 1] - h.level_start[L] + 31) / 32;
            trips += trips & 1;
            const int cap = 32 * (trips + 1);
            const unsigned spare = (unsigned)h.n_pix * (unsigned)sizeof(T); // byte offset of the always-zero cell behind the image
            std::vector<unsigned char> img((size_t)cap * (sizeof(W4<T>) + sizeof(uint2) + sizeof(unsigned short)));
            W4<T> *tw = reinterpret_cast<W4<T> *>(img.data());
            uint2 *tn = reinterpret_cast<uint2 *>(tw + cap);
            unsigned short *tp = reinterpret_cast<unsigned short *>(tn + cap);
            for (int q = 0; q < cap; ++q) {
                tw[q] = W4<T>{T(0), T(0), T(0), T(0)};
                tn[q] = make_uint2(spare | (spare << 16), spare | (spare << 16));
                tp[q] = (unsigned short)spare;
            }
            int t = 0;
            for (int L = 0; L < h.n_levels; ++L)
                for (int b = h.level_start[L]; b < h.level_start[L + 1]; b += 32, ++t)
                    for (int l = 0; l < 32 && b + l < h.level_start[L + 1]; ++l) {
                        const int j = b + l, q = 32 * t + l;
                        unsigned o[4];
                        for (int i = 0; i < 4; ++i) {
                            const unsigned v = h.nbr[(size_t)i * n + j];
                            o[i] = v == 0xffffu ? spare : v * (unsigned)sizeof(T);
                        }
                        tn[q] = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
                        tw[q] = W4<T>{(T)h.w[j], (T)h.w[(size_t)n + j], (T)h.w[(size_t)2 * n + j], (T)h.w[(size_t)3 * n + j]};
                        tp[q] = (unsigned short)((unsigned)h.pix[j] * (unsigned)sizeof(T));
                    }
            SB_TRY(wtab.alloc(img.size()));
            SB_CUDA(cudaMemcpy(wtab.p, img.data(), img.size(), cudaMemcpyHostToDevice));
            dev.wtab = wtab.p, dev.w_trips = trips, dev.w_cap = cap;
        }
        return SB_OK;
    }
};

static int check_chain(const sb_chain_desc &c, int n_mono) {
    if (c.n_ops < 0 || c.n_ops > SB_MAX_CHAIN_OPS || c.repeat < 1) return set_err(SB_ERR_ARG, "bad constraint chain");
    for (int i = 0; i < c.n_ops; ++i) {
        const sb_op &o = c.ops[i];
        if (o.code < SB_OP_MONOTONIC || o.code > SB_OP_NORMALIZE) return set_err(SB_ERR_ARG, "unknown constraint op-code %d", o.code);
        if (o.code == SB_OP_MONOTONIC && (o.iarg < 0 || o.iarg >= n_mono)) return set_err(SB_ERR_ARG, "monotonic table index %d out of range", o.iarg);
    }
    return SB_OK;
}

// The dynamic shared-memory limit of a kernel is a per-device function attribute shared by every plan of the process:
// only ever raise it (a second, smaller plan must not lower the limit under a plan that is still alive).
static std::mutex g_smem_mutex;
static std::map<std::pair<int, const void *>, size_t> g_smem_limit;
static int raise_smem_limit(int device, const void *fn, size_t bytes) {
    std::lock_guard<std::mutex> lock(g_smem_mutex);
    size_t &cur = g_smem_limit[std::make_pair(device, fn)];
    if (bytes > cur) {
        SB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
    return SB_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
typedef CUresult (*tensor_map_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tensor_map_encode_fn tensor_map_encoder() {
    static tensor_map_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tensor_map_encode_fn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
// 2-D map over a row-major [rows][pitch] array of 8-byte elements (complex64), box = [box_rows][box_cols]
static bool make_tensor_map_c64(CUtensorMap *map, void *base, uint64_t rows, uint64_t pitch, uint32_t box_rows, uint32_t box_cols) {
    tensor_map_encode_fn enc = tensor_map_encoder();
    if (!enc || box_rows > 256 || box_cols > 256 || (pitch * 8) % 16 != 0) return false;
    const cuuint64_t dims[2] = {pitch, rows}, strides[1] = {pitch * 8};
    const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
    long long g = (n + block - 1) / block;
    return (int)std::max<long long>(1, std::min<long long>(g, cap));
}

} // namespace sb

using namespace sb;

// =====================================================================================================
// plan
// =====================================================================================================
struct sb_plan {
    virtual ~sb_plan() {}
    virtual int upload_observation(int obs, const void *data, const void *weights, int elem_bytes, const double *khat, const double *loss_const) = 0;
    virtual int upload_kernels(int obs, const double *ker, int Py, int Px, int y0, int x0) = 0;
    virtual int zero_state() = 0;
    virtual int upload_resampling(int obs, const double *ey, const double *ex, double h2) = 0;
    virtual int upload_resampling_rot(int obs, const double *ra, const double *rb, double h2) = 0;
    virtual int upload_params(int which, const double *sed, const double *morph, const double *center) = 0;
    virtual int download_params(int which, double *sed, double *morph, double *center) = 0;
    virtual int evaluate(int obs, double *model, double *rendered, double *loss, double *g_sed, double *g_morph, double *g_center) = 0;
    virtual int fit(const sb_fit_opts *o, int32_t *n_iter, double *loss, int32_t *status) = 0;
    virtual int fit_enqueue(const sb_fit_opts *o, int n) = 0;
    virtual int profile(const sb_fit_opts *o, int n, float *stage_ms) = 0;
    virtual int device_params(void **sed, int64_t *n_sed, void **morph, int64_t *n_morph, int *elem_bytes) = 0;
    virtual int spectral_mode() const = 0;
    virtual int prox_histogram(int enable, int64_t *out16) = 0;
    virtual int scene_control(const int32_t *it_local, const int32_t *loss_len, const int32_t *limit, const int32_t *active, const int32_t *prox_iter) = 0;
    virtual int scene_status(int32_t *it_local, int32_t *loss_len, int32_t *state) = 0;
    virtual int run(const sb_fit_opts *o, int max_launches, int32_t *launched) = 0;
    virtual int download_loss(double *loss, int n_cols) = 0;
    virtual int upload_loss(const double *loss, int n_cols) = 0;
    virtual int inspect(int32_t *action) = 0;
    virtual int set_sources(const sb_batch_desc *desc) = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t side = nullptr; // runs the generic update kernel next to the grouped one
    int64_t dev_bytes = 0;
    int64_t launches = 0;
};

template <typename T> struct PlanT : sb_plan {
    typedef typename Cx<T>::type cplx;
    sb_batch_desc desc;
    int S = 0, C = 0, n_src = 0, n_point = 0, npix_max = 0, npix_shift = 0;
    long long n_morph = 0, n_pmorph = 0;
    std::vector<DevSource> h_src;
    std::vector<int> h_start;
    DevBuf<DevSource> d_src;
    DevBuf<int> d_start, d_done, d_niter, d_status, d_it, d_nactive, d_nactive_next, d_state, d_limit, d_prox_iter;
    bool use_limit = false, use_prox_iter = false;
    DevBuf<double> d_sed, d_sed_m, d_sed_v, d_sed_vhat, d_center, d_cen_m, d_cen_v, d_cen_vhat, d_loss, d_loss_const;
    DevBuf<double> d_gsed, d_gcenter, d_gmorph, d_stage;
    DevBuf<T> d_morph, d_morph_m, d_morph_v, d_morph_vhat, d_pmorph, d_model, d_rendered, d_scratch_x, d_scratch_ps;
    DevBuf<int> d_work, d_fast_groups, d_shift_list;
    DevBuf<T> d_smorph, d_toep;
    int n_shift = 0, toep_len = 1;
    int n_generic = 0, n_fast_cta = 0, fast_G = 0, fast_GT = 64, fast_npix = 0, fast_table_cap = 0;
    size_t fast_smem = 0;
    // warp-per-source kernel (update_warp.cuh)
    static constexpr int WARP_NPT = 56, WARP_MAXT = sizeof(T) == 4 ? 448 : 256;
    static constexpr size_t WARP_RING = XpRing<T>::BYTES;
    DevBuf<int> d_warp_groups, d_warp_ctr;
    DevBuf<int4> d_warp_segs;
    int n_warp_chains = 0;
    DevBuf<XP<T>> d_xp;
    int n_warp_cta = 0, warp_G = 0, warp_npix = 0, warp_cap = 0;
    size_t warp_smem = 0;
    cudaStream_t side2 = nullptr;
    cudaEvent_t ev_join2 = nullptr;
    std::vector<HostMono> hmonos;
    DevBuf<float> d_stage_f;
    DevBuf<DevChain> d_chains;
    DevBuf<DevMono> d_monos;
    std::vector<std::unique_ptr<MonoDevice<T>>> monos;
    struct Obs {
        DevObs<T> dev;
        DevBuf<T> A, B, data, weights;
        DevBuf<double> partials;
        int n_part = 0;
        DevBuf<cplx> Ahat, khat;
        cufftHandle fwd = 0, inv = 0;
        bool have_plans = false;
        std::vector<double> loss_const;
        // fused spectral path (spectral.cuh)
        bool fused = false;
        SpecObs<T> sdev;
        SpecKernels<T> kx, ky;
        DevBuf<cplx> X, tw_x, tw_y, Pbuf, T1buf, Ey, Ex; // Pbuf..Ex: resampling observations (kind 2) only
        DevBuf<int> cand_start, cand;                     // render kernel: sources per (scene, row block)
        int max_cand = 0;
        DevBuf<cplx> RA, RB;                              // rotated resampling observations (kind 3): multiplier tables
        DevBuf<T> Rres, Rpart;                            //   weighted residual, per-chunk partial renders
        int rot_chunks = 0, rot_chunk = 0, rot_cb = 1;
        size_t rot_smem = 0;
        DevBuf<T> G;
        int npair = 1, cb = 1, row_threads = 0;
        bool render_two = false;
        size_t smem_render = 0, smem_row = 0, smem_col = 0, smem_col_tma = 0;
        bool use_tma = false; // column pass through the Tensor Memory Accelerator (float, Ny <= 256)
        CUtensorMap tmX, tmK;
        // ConvolutionRenderer(psf_shift=...): fitted kernel offset (psf_shift.cuh)
        bool psf_shift = false, have_kernel = false;
        int slot0 = 0, Py = 0, Px = 0, ky0 = 0, kx0 = 0, shift_Fy = 0, shift_Fx = 0, shift_fixed = 0;
        double shift_step = 0;
        DevBuf<double> ks, kd0, kd1, gk;
        DevBuf<T> resid;
        // kernel-image -> K^ (double precision, chunked over scenes)
        cufftHandle kplan = 0;
        bool have_kplan = false;
        int kchunk = 0;
        DevBuf<double> kimg, kgrid;
        DevBuf<double2> kspec;
        // upload staging, kept between calls: cudaFree synchronises the whole device, which would make a copy-in wait for
        // another plan's running loop (BatchPipeline)
        DevBuf<float> stage_f;
        DevBuf<double2> stage_z;
    };
    std::vector<std::unique_ptr<Obs>> obs;
    int loss_cap = 0;
    int *h_nactive = nullptr; // pinned
    cudaGraphExec_t graph = nullptr;
    FitScalars graph_fs;
    bool have_graph = false;
    int kernels_per_iter = 0, ffts_per_iter = 0;
    bool use_fast = getenv("SB_NO_FAST_UPDATE") == nullptr;
    bool fused = false; // every observation runs the fused spectral kernels (else: cuFFT path for all)
    int max_src_scene = 0;

    ~PlanT() override {
        if (graph) cudaGraphExecDestroy(graph);
        for (auto &o : obs) {
            if (o->have_plans) {
                cufftDestroy(o->fwd);
                cufftDestroy(o->inv);
            }
            if (o->have_kplan) cufftDestroy(o->kplan);
        }
        if (h_nactive) cudaFreeHost(h_nactive);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_join2) cudaEventDestroy(ev_join2);
        if (stream) cudaStreamDestroy(stream);
        if (side) cudaStreamDestroy(side);
        if (side2) cudaStreamDestroy(side2);
    }

    int64_t total_bytes() {
        int64_t b = d_src.bytes() + d_sed.bytes() * 4 + d_center.bytes() * 4 + d_loss.bytes() + d_morph.bytes() * 4 + d_pmorph.bytes() +
                    d_model.bytes() + d_rendered.bytes() + d_gmorph.bytes();
        for (auto &o : obs)
            b += o->A.bytes() + o->B.bytes() + o->data.bytes() + o->weights.bytes() + o->Ahat.bytes() + o->khat.bytes() + o->X.bytes() +
                 o->G.bytes();
        return b;
    }

    int init(const sb_batch_desc *dsc, int dev) {
        desc = *dsc;
        device = dev;
        S = desc.n_scenes, C = desc.C, n_src = desc.n_sources;
        if (S <= 0 || C <= 0 || C > SB_MAXC || desc.Ny <= 0 || desc.Nx <= 0) return set_err(SB_ERR_ARG, "bad frame (S=%d C=%d Ny=%d Nx=%d)", S, C, desc.Ny, desc.Nx);
        if (desc.n_obs <= 0 || desc.n_obs > SB_MAX_OBS) return set_err(SB_ERR_ARG, "n_obs=%d out of range", desc.n_obs);
        if (n_src < 0 || !desc.scene_src_start || (n_src && !desc.sources)) return set_err(SB_ERR_ARG, "missing source tables");
        SB_CUDA(cudaSetDevice(device));
        SB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        SB_CUDA(cudaEventCreate(&ev0));
        SB_CUDA(cudaEventCreate(&ev1));
        SB_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        SB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        SB_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        SB_CUDA(cudaStreamCreateWithFlags(&side2, cudaStreamNonBlocking));
        SB_CUDA(cudaEventCreateWithFlags(&ev_join2, cudaEventDisableTiming));
        SB_CUDA(cudaHostAlloc((void **)&h_nactive, sizeof(int), cudaHostAllocDefault));
        SB_TRY(init_sources());
        SB_TRY(init_observations());
        return SB_OK;
    }

    // Everything that depends on the sources (boxes, chains, wavefront tables, parameter arrays, update-kernel groups).  Runs at
    // plan creation and again from sb_plan_set_sources (dynamic boxes): every buffer is (re)allocated, nothing is carried over.
    std::vector<int> psf_slot0;
    int init_sources() {
        n_src = desc.n_sources;
        if (n_src < 0 || !desc.scene_src_start || (n_src && !desc.sources)) return set_err(SB_ERR_ARG, "missing source tables");
        n_point = 0, npix_max = 0, npix_shift = 0, n_morph = 0, n_pmorph = 0, n_shift = 0, toep_len = 1;
        monos.clear();
        hmonos.clear();
        // ---- constraint tables
        for (int i = 0; i < desc.n_chains; ++i) SB_TRY(check_chain(desc.chains[i], desc.n_mono));
        std::vector<DevMono> hm(std::max(desc.n_mono, 1));
        for (int i = 0; i < desc.n_mono; ++i) {
            HostMono h;
            SB_TRY(build_mono<T>(desc.mono[i], h));
            monos.emplace_back(new MonoDevice<T>());
            SB_TRY(monos.back()->upload(h, stream));
            hm[i] = monos.back()->dev;
            h.nbr.clear(), h.w.clear(), h.pix.clear();
            hmonos.push_back(h);
        }
        SB_TRY(d_monos.alloc(hm.size()));
        SB_CUDA(cudaMemcpy(d_monos.p, hm.data(), hm.size() * sizeof(DevMono), cudaMemcpyHostToDevice));
        std::vector<DevChain> hc(std::max(desc.n_chains, 1));
        for (int i = 0; i < desc.n_chains; ++i) {
            hc[i].n_ops = desc.chains[i].n_ops, hc[i].repeat = desc.chains[i].repeat;
            memcpy(hc[i].ops, desc.chains[i].ops, sizeof(hc[i].ops));
        }
        SB_TRY(d_chains.alloc(hc.size()));
        SB_CUDA(cudaMemcpy(d_chains.p, hc.data(), hc.size() * sizeof(DevChain), cudaMemcpyHostToDevice));

        // ---- sources
        h_start.assign(desc.scene_src_start, desc.scene_src_start + S + 1);
        if (h_start[0] != 0 || h_start[S] != n_src) return set_err(SB_ERR_ARG, "scene_src_start must run from 0 to n_sources");
        h_src.resize(n_src);
        const int b = desc.psf_boxsize;
        for (int s = 0; s < S; ++s) {
            if (h_start[s + 1] < h_start[s]) return set_err(SB_ERR_ARG, "scene_src_start not monotone");
            for (int k = h_start[s]; k < h_start[s + 1]; ++k) {
                const sb_source_desc &in = desc.sources[k];
                DevSource &d = h_src[k];
                memset(&d, 0, sizeof d);
                d.kind = in.kind, d.By = in.By, d.Bx = in.Bx, d.oy = in.oy, d.ox = in.ox, d.chain = in.chain, d.sed_chain = in.sed_chain;
                d.sed_is_f32 = in.sed_is_f32, d.morph_fixed = in.morph_fixed, d.sed_fixed = in.sed_fixed, d.scene = s;
                d.morph_step = in.morph_step, d.sed_step_factor = in.sed_step_factor, d.resizing = in.kind == 0 ? in.resizing : 0;
                memcpy(d.sed_step_min, in.sed_step_min, sizeof d.sed_step_min);
                if (d.By <= 0 || d.Bx <= 0 || (long long)d.By * d.Bx >= 0xffff) return set_err(SB_ERR_ARG, "source %d: bad box %dx%d", k, d.By, d.Bx);
                if (d.chain >= desc.n_chains || d.sed_chain >= desc.n_chains) return set_err(SB_ERR_ARG, "source %d: chain index out of range", k);
                if (d.sed_chain >= 0)
                    for (int i = 0; i < desc.chains[d.sed_chain].n_ops; ++i) {
                        const int code = desc.chains[d.sed_chain].ops[i].code;
                        if (code != SB_OP_POSITIVITY && code != SB_OP_NORMALIZE)
                            return set_err(SB_ERR_ARG, "source %d: spectrum constraint op-code %d is not supported on a 1-D parameter", k, code);
                    }
                if (d.kind == 0 && in.shifting) {
                    if (in.shift_Fy < in.By || in.shift_Fx < in.Bx || (in.shift_Fx & 1))
                        return set_err(SB_ERR_ARG, "source %d: bad shift grid %dx%d for a %dx%d image", k, in.shift_Fy, in.shift_Fx, in.By, in.Bx);
                    d.shifting = 1, d.shift_Fy = in.shift_Fy, d.shift_Fx = in.shift_Fx, d.shift_step = in.shift_step;
                    d.toep_off = n_shift++;
                    toep_len = std::max(toep_len, 2 * std::max(in.By, in.Bx) - 1);
                    npix_shift = std::max(npix_shift, (in.By * in.Bx + 3) & ~3);
                }
                if (d.kind == 0) {
                    if (d.chain >= 0)
                        for (int i = 0; i < desc.chains[d.chain].n_ops; ++i) {
                            const sb_op &op = desc.chains[d.chain].ops[i];
                            if (op.code == SB_OP_MONOTONIC && desc.mono[op.iarg].n_pix != d.By * d.Bx)
                                return set_err(SB_ERR_ARG, "source %d: monotonic table size %d != box %dx%d", k, desc.mono[op.iarg].n_pix, d.By, d.Bx);
                        }
                    d.morph_off = n_morph;
                    d.point_idx = d.shifting ? n_point++ : -1;
                    n_morph += (long long)d.By * d.Bx;
                    npix_max = std::max(npix_max, d.By * d.Bx);
                } else if (d.kind == 1) {
                    if (b <= 0 || b > 15 || (b & 1) == 0) return set_err(SB_ERR_ARG, "point source needs an odd model-PSF box <= 15 (got %d)", b);
                    if (d.By != b || d.Bx != b) return set_err(SB_ERR_ARG, "source %d: point-source box must equal the model PSF box", k);
                    d.morph_off = n_pmorph;
                    d.point_idx = n_point++;
                    n_pmorph += (long long)C * b * b;
                } else
                    return set_err(SB_ERR_ARG, "source %d: unknown kind %d", k, d.kind);
            }
        }
        npix_max = (npix_max + 3) & ~3;
        // renderer parameters (psf_shift) take centre slots behind the sources': observation-major, one per scene
        psf_slot0.assign(desc.n_obs, -1);
        for (int o = 0; o < desc.n_obs; ++o)
            if (desc.obs[o].psf_shift) {
                psf_slot0[o] = n_point;
                n_point += S;
            }
        SB_TRY(plan_fast_path());
        SB_TRY(d_src.alloc(std::max(n_src, 1)));
        SB_TRY(d_start.alloc(S + 1));
        if (n_src) SB_CUDA(cudaMemcpy(d_src.p, h_src.data(), n_src * sizeof(DevSource), cudaMemcpyHostToDevice));
        SB_CUDA(cudaMemcpy(d_start.p, h_start.data(), (S + 1) * sizeof(int), cudaMemcpyHostToDevice));
        const size_t nsed = (size_t)std::max(n_src, 1) * C, ncen = (size_t)std::max(n_point, 1) * 2;
        DevBuf<double> *dz[] = {&d_sed, &d_sed_m, &d_sed_v, &d_sed_vhat, &d_gsed};
        for (auto *bf : dz) {
            SB_TRY(bf->alloc(nsed));
            SB_TRY(bf->zero(stream));
        }
        DevBuf<double> *dc[] = {&d_center, &d_cen_m, &d_cen_v, &d_cen_vhat, &d_gcenter};
        for (auto *bf : dc) {
            SB_TRY(bf->alloc(ncen));
            SB_TRY(bf->zero(stream));
        }
        DevBuf<T> *dm[] = {&d_morph, &d_morph_m, &d_morph_v, &d_morph_vhat};
        for (auto *bf : dm) {
            SB_TRY(bf->alloc(std::max<long long>(n_morph, 1)));
            SB_TRY(bf->zero(stream));
        }
        SB_TRY(d_pmorph.alloc(std::max<long long>(n_pmorph, 1)));
        SB_TRY(d_pmorph.zero(stream));
        if (n_shift) {
            std::vector<int> sl;
            for (int k = 0; k < n_src; ++k)
                if (h_src[k].shifting) sl.push_back(k);
            SB_TRY(d_shift_list.alloc(sl.size()));
            SB_CUDA(cudaMemcpy(d_shift_list.p, sl.data(), sl.size() * sizeof(int), cudaMemcpyHostToDevice));
            SB_TRY(d_smorph.alloc(std::max<long long>(n_morph, 1)));
            SB_TRY(d_smorph.zero(stream));
            SB_TRY(d_toep.alloc((size_t)n_shift * 8 * toep_len));
            SB_TRY(d_toep.zero(stream));
            if ((size_t)3 * npix_shift * sizeof(T) > 220 * 1024)
                return set_err(SB_ERR_ARG, "largest box of a shifting source (%d px) does not fit in shared memory", npix_shift);
            SB_TRY(raise_smem((const void *)k_shift_apply<T>, (size_t)3 * npix_shift * sizeof(T)));
        }
        if (!d_done.p) { // per-scene run state and loss constants: allocated once, they outlive a change of sources
            SB_TRY(d_done.alloc(S));
            SB_TRY(d_niter.alloc(S));
            SB_TRY(d_status.alloc(S));
            SB_TRY(d_it.alloc(S));
            SB_TRY(d_state.alloc(S));
            SB_TRY(d_limit.alloc(S));
            SB_TRY(d_prox_iter.alloc(S));
            SB_TRY(d_nactive.alloc(1));
            SB_TRY(d_nactive_next.alloc(1));
            SB_TRY(d_loss_const.alloc(S));
            SB_TRY(d_loss_const.zero(stream));
        }
        SB_TRY(ensure_loss_cap(256));
        max_src_scene = 0;
        for (int s_ = 0; s_ < S; ++s_) max_src_scene = std::max(max_src_scene, h_start[s_ + 1] - h_start[s_]);
        // dynamic shared memory of the generic update kernel
        const size_t smem = update_smem();
        if (smem > 227 * 1024) return set_err(SB_ERR_ARG, "largest morphology box (%d px) does not fit in shared memory", npix_max);
        SB_TRY(raise_smem((const void *)k_update<T>, smem));
        return SB_OK;
    }

    static size_t resample_lr_smem(int H, int W, int Fxc) { // spectra [H][Fxc], shift table [W][Fxc], residual [H][W]
        return (size_t)(H + W) * Fxc * sizeof(cplx) + (size_t)H * W * sizeof(T);
    }
    size_t render_smem(const Obs &ob) const {
        return ob.smem_row + (size_t)ob.cb * 2 * ob.npair * desc.Nx * sizeof(T) + 16 + (size_t)(ob.max_cand + 1) * sizeof(SpecCand<T>);
    }
    // Render kernel: per (scene, block of 2 npair frame rows) the sources whose boxes touch the rows, in scene order.  Boxes are
    // fixed for the life of the source tables (sb_plan_set_sources rebuilds the lists).
    int build_candidates(Obs &ob) {
        const int rows = 2 * ob.npair, nblk = (desc.Ny + rows - 1) / rows;
        std::vector<int> start((size_t)S * nblk + 1, 0), list;
        ob.max_cand = 0;
        int max_npx = 0; // most pixels of one source inside one row block
        for (int s = 0; s < S; ++s)
            for (int b = 0; b < nblk; ++b) {
                const int y0 = b * rows;
                start[(size_t)s * nblk + b] = (int)list.size();
                // bit 31 of an entry: the source's columns overlap those of an earlier source added since the last barrier --
                // the kernel then waits before adding it; sources side by side need no barrier between them
                std::vector<std::pair<int, int>> group;
                for (int k = h_start[s]; k < h_start[s + 1]; ++k)
                    if (h_src[k].oy < y0 + rows && h_src[k].oy + h_src[k].By > y0) {
                        const int x0 = h_src[k].ox, x1 = h_src[k].ox + h_src[k].Bx;
                        const int nr = std::min(std::min(y0 + rows, desc.Ny), h_src[k].oy + h_src[k].By) - std::max(y0, h_src[k].oy);
                        max_npx = std::max(max_npx, nr * (std::min(x1, desc.Nx) - std::max(x0, 0)));
                        bool hit = false;
                        for (auto &g : group) hit = hit || (x0 < g.second && g.first < x1);
                        if (hit) group.clear();
                        group.emplace_back(x0, x1);
                        list.push_back(k | (hit ? (int)0x80000000u : 0));
                    }
                ob.max_cand = std::max(ob.max_cand, (int)list.size() - start[(size_t)s * nblk + b]);
            }
        start[(size_t)S * nblk] = (int)list.size();
        // the render kernel adds a source with one thread per covered pixel and requests the pixels of four sources ahead: where
        // a box has more pixels in a row block than the CTA has threads, the two-pixels-per-thread instantiation is used
        ob.render_two = max_npx > ob.row_threads && max_npx <= 2 * ob.row_threads;
        if (ob.cand_start.n < start.size()) SB_TRY(ob.cand_start.alloc(start.size()));
        if (ob.cand.n < std::max<size_t>(list.size(), 1)) SB_TRY(ob.cand.alloc(std::max<size_t>(list.size(), 1) * 5 / 4 + 16));
        SB_CUDA(cudaMemcpyAsync(ob.cand_start.p, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        if (!list.empty()) SB_CUDA(cudaMemcpyAsync(ob.cand.p, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        SB_CUDA(cudaStreamSynchronize(stream)); // the host vectors go out of scope
        ob.sdev.cand_start = ob.cand_start.p, ob.sdev.cand = ob.cand.p;
        return SB_OK;
    }

    // sb_plan_set_sources: new boxes / chains / tables for the same scenes and observations
    DevBuf<int> d_action;
    int inspect(int32_t *action) override {
        SB_CUDA(cudaSetDevice(device));
        if (!action) return set_err(SB_ERR_ARG, "null action array");
        if (d_action.n < (size_t)std::max(n_src, 1)) SB_TRY(d_action.alloc(std::max(n_src, 1)));
        if (n_src) {
            k_inspect<T><<<n_src, 128, 0, stream>>>(d_src.p, n_src, d_morph.p, d_morph_m.p, d_morph_v.p, d_state.p, d_action.p);
            SB_CUDA(cudaGetLastError());
            SB_CUDA(cudaMemcpyAsync(action, d_action.p, (size_t)n_src * sizeof(int), cudaMemcpyDeviceToHost, stream));
        }
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }

    int set_sources(const sb_batch_desc *dsc) override {
        SB_CUDA(cudaSetDevice(device));
        SB_CUDA(cudaStreamSynchronize(stream));
        if (dsc->n_scenes != S || dsc->C != C || dsc->Ny != desc.Ny || dsc->Nx != desc.Nx || dsc->n_obs != desc.n_obs ||
            dsc->precision != desc.precision || memcmp(dsc->obs, desc.obs, sizeof desc.obs) != 0)
            return set_err(SB_ERR_ARG, "sb_plan_set_sources: scenes, frame and observations must be the plan's");
        desc = *dsc;
        SB_TRY(init_sources());
        for (size_t o = 0; o < obs.size(); ++o) {
            Obs &ob = *obs[o];
            if (ob.psf_shift) ob.slot0 = psf_slot0[o];
            if (ob.fused) {
                SB_TRY(build_candidates(ob));
                ob.smem_render = render_smem(ob);
                if (ob.smem_render > 227 * 1024) return set_err(SB_ERR_ARG, "observation %d: frame too wide for the fused spectral kernels", (int)o);
                SB_TRY(raise_smem((const void *)ob.kx.render, ob.smem_render));
                SB_TRY(raise_smem((const void *)ob.kx.render2, ob.smem_render));
            }
        }
        have_graph = false;
        drop_scene_tables();
        SB_TRY(reset_counters());
        SB_CUDA(cudaStreamSynchronize(stream));
        dev_bytes = total_bytes();
        return SB_OK;
    }

    int init_observations() {
        // ---- observations
        {
            const char *mode = getenv("SB_SPECTRAL");
            fused = !(mode && strcmp(mode, "cufft") == 0);
            for (int o = 0; o < desc.n_obs && fused; ++o) {
                const sb_obs_desc &od = desc.obs[o];
                SpecKernels<T> k;
                if ((od.kind != 0 && od.kind != 2 && od.kind != 3) || !spec_kernels<T>(od.Fx, &k) || !spec_kernels<T>(od.Fy, &k)) fused = false;
            }
        }
        for (int o = 0; o < desc.n_obs; ++o) {
            const sb_obs_desc &od = desc.obs[o];
            obs.emplace_back(new Obs());
            Obs &ob = *obs.back();
            if (od.C <= 0 || od.chan_off < 0 || od.chan_off + od.C > C) return set_err(SB_ERR_ARG, "observation %d: channels outside the model frame", o);
            if (od.kind < 0 || od.kind > 3) return set_err(SB_ERR_ARG, "observation %d: unsupported renderer kind %d", o, od.kind);
            if (od.psf_shift && !fused)
                return set_err(SB_ERR_ARG, "observation %d: psf_shift needs FFT lengths of the fused spectral kernels (got %dx%d)", o, od.Fy, od.Fx);
            if ((od.kind == 2 || od.kind == 3) && !fused)
                return set_err(SB_ERR_ARG, "observation %d: a resampling observation needs FFT lengths of the fused spectral kernels "
                                           "(got %dx%d) and no NullRenderer observation in the same plan", o, od.Fy, od.Fx);
            int Fy = od.Fy, Fx = od.Fx;
            if (od.kind == 1) Fy = desc.Ny, Fx = desc.Nx;
            if (Fy < desc.Ny || Fx < desc.Nx) return set_err(SB_ERR_ARG, "observation %d: FFT grid %dx%d smaller than the frame", o, Fy, Fx);
            DevObs<T> &d = ob.dev;
            memset(&d, 0, sizeof d);
            d.kind = od.kind, d.C = od.C, d.H = od.H, d.W = od.W, d.chan_off = od.chan_off, d.oy = od.oy, d.ox = od.ox;
            d.Fy = Fy, d.Fx = Fx, d.Fxc = Fx / 2 + 1, d.khat_shared = od.khat_shared;
            const size_t ndata = (size_t)S * od.C * od.H * od.W;
            SB_TRY(ob.data.alloc(ndata));
            SB_TRY(ob.weights.alloc(ndata));
            SB_TRY(ob.data.zero(stream));
            SB_TRY(ob.weights.zero(stream));
            ob.loss_const.assign(S, 0.0);
            ob.fused = fused;
            if (fused) {
                spec_kernels<T>(Fx, &ob.kx);
                spec_kernels<T>(Fy, &ob.ky);
                const int Xp = (d.Fxc + (sizeof(T) == 4 ? 3 : 1)) & ~(sizeof(T) == 4 ? 3 : 1);
                d.Kp = Xp, d.Bh = desc.Ny, d.Bw = desc.Nx;
                SB_TRY(ob.X.alloc((size_t)S * od.C * desc.Ny * Xp));
                SB_TRY(ob.X.zero(stream));
                SB_TRY(ob.G.alloc((size_t)S * od.C * desc.Ny * desc.Nx));
                SB_TRY(ob.G.zero(stream));
                SB_TRY(ob.khat.alloc((size_t)(od.khat_shared ? 1 : S) * od.C * Fy * Xp));
                SB_TRY(ob.khat.zero(stream));
                SB_TRY(upload_twiddles(ob.tw_x, ob.kx.R1, ob.kx.R2));
                SB_TRY(upload_twiddles(ob.tw_y, ob.ky.R1, ob.ky.R2));
                // launch geometry of the row kernels: cb bands x npair row pairs per CTA
                const int rmax = std::max(ob.kx.R1, ob.kx.R2), limit = sizeof(T) == 4 ? 512 : 256;
                ob.cb = std::min(od.C, 8);
                while (ob.cb > 1 && ob.cb * rmax > limit) --ob.cb;
                ob.npair = std::max(1, std::min(4, 224 / (ob.cb * rmax)));
                ob.row_threads = ((ob.npair * ob.cb * rmax + 31) / 32) * 32;
                if (ob.row_threads > limit) ob.row_threads = ob.npair * ob.cb * rmax;
                const size_t nb = (size_t)ob.npair * ob.cb;
                ob.smem_row = (nb * ob.kx.sf + (size_t)Fx) * sizeof(cplx) + 16;
                SB_TRY(build_candidates(ob));
                ob.smem_render = render_smem(ob);
                ob.smem_col = ((size_t)ob.ky.NBcol * ob.ky.sf + (size_t)Fy) * sizeof(cplx) + 16;
                if (ob.smem_render > 227 * 1024 || ob.smem_col > 227 * 1024)
                    return set_err(SB_ERR_ARG, "observation %d: frame too wide for the fused spectral kernels", o);
                SB_TRY(raise_smem((const void *)ob.kx.render, ob.smem_render));
                SB_TRY(raise_smem((const void *)ob.kx.render2, ob.smem_render));
                SB_TRY(raise_smem((const void *)ob.kx.residual, ob.smem_row));
                SB_TRY(raise_smem((const void *)ob.kx.residual_r, ob.smem_row));
                SB_TRY(raise_smem((const void *)ob.kx.grad, ob.smem_row));
                SB_TRY(raise_smem((const void *)ob.ky.column, ob.smem_col));
                if (sizeof(T) == 4 && od.kind == 0 && ob.ky.column_tma && desc.Ny <= 256 && Fy % 2 == 0 && Fy / 2 <= 256 &&
                    getenv("SB_NO_TMA") == nullptr) {
                    const int NB = ob.ky.NBcol;
                    ob.smem_col_tma = 128 + ((size_t)(desc.Ny + Fy) * NB + Fy) * sizeof(cplx) + 16;
                    ob.use_tma = ob.smem_col_tma <= 227 * 1024 &&
                                 make_tensor_map_c64(&ob.tmX, ob.X.p, (uint64_t)S * od.C * desc.Ny, Xp, desc.Ny, NB) &&
                                 make_tensor_map_c64(&ob.tmK, ob.khat.p, (uint64_t)(od.khat_shared ? 1 : S) * od.C * Fy, Xp, Fy / 2, NB);
                    if (ob.use_tma) SB_TRY(raise_smem((const void *)ob.ky.column_tma, ob.smem_col_tma));
                }
                ob.n_part = ((desc.Ny + 2 * ob.npair - 1) / (2 * ob.npair)) * ((od.C + ob.cb - 1) / ob.cb);
                if (od.kind == 2) {
                    ob.n_part = od.C;
                    SB_TRY(ob.Pbuf.alloc((size_t)S * od.C * Fy * Xp));
                    SB_TRY(ob.T1buf.alloc((size_t)S * od.C * od.H * Xp));
                    SB_TRY(ob.Ey.alloc((size_t)od.H * Fy));
                    SB_TRY(ob.Ex.alloc((size_t)od.W * d.Fxc));
                    SB_TRY(ob.Pbuf.zero(stream));
                    SB_TRY(ob.T1buf.zero(stream));
                    SB_TRY(ob.Ey.zero(stream));
                    SB_TRY(ob.Ex.zero(stream));
                    if (resample_lr_smem(od.H, od.W, d.Fxc) > 200 * 1024) return set_err(SB_ERR_ARG, "observation %d: resampled image too large", o);
                    SB_TRY(raise_smem((const void *)ob.ky.column_fwd, ob.smem_col));
                    SB_TRY(raise_smem((const void *)ob.ky.column_inv, ob.smem_col));
                    SB_TRY(raise_smem((const void *)k_resample_lr<T>, resample_lr_smem(od.H, od.W, d.Fxc)));
                    if ((size_t)Fy * SB_RS_ROWS * sizeof(cplx) > 200 * 1024 || (size_t)od.H * SB_RS_KY * sizeof(cplx) > 200 * 1024)
                        return set_err(SB_ERR_ARG, "observation %d: resampling geometry too large", o);
                    SB_TRY(raise_smem((const void *)k_resample_t1<T>, (size_t)Fy * SB_RS_ROWS * sizeof(cplx)));
                    SB_TRY(raise_smem((const void *)k_resample_q<T>, (size_t)od.H * SB_RS_KY * sizeof(cplx)));
                }
                if (od.kind == 3) {
                    if (od.H > 32 || od.W > 32) return set_err(SB_ERR_ARG, "observation %d: a rotated resampling observation is limited to 32x32 pixels (got %dx%d)", o, od.H, od.W);
                    const size_t K = (size_t)Fy * Xp;
                    // k range per CTA: at least two CTAs per SM in flight when the batch is small, 2048 entries at most
                    ob.rot_cb = SB_ROT_MAXB;
                    const int groups = (od.C + ob.rot_cb - 1) / ob.rot_cb;
                    long long want = (2 * 148 + (long long)S * groups - 1) / ((long long)S * groups);
                    want = std::max<long long>(want, (long long)(K + 2047) / 2048);
                    want = std::min<long long>(want, (long long)(K + 255) / 256);
                    ob.rot_chunk = (int)(((K + want - 1) / want + SB_ROT_SUB - 1) / SB_ROT_SUB) * SB_ROT_SUB;
                    ob.rot_chunks = (int)((K + ob.rot_chunk - 1) / ob.rot_chunk);
                    ob.rot_smem = (size_t)(2 + 2 * ob.rot_cb) * SB_ROT_SUB * SB_ROT_LD * sizeof(T) + (size_t)2 * ob.rot_cb * SB_ROT_SUB * sizeof(cplx);
                    ob.n_part = od.C;
                    SB_TRY(ob.Pbuf.alloc((size_t)S * od.C * K));
                    SB_TRY(ob.RA.alloc((size_t)od.H * K));
                    SB_TRY(ob.RB.alloc((size_t)od.W * K));
                    SB_TRY(ob.Rres.alloc(ndata));
                    SB_TRY(ob.Rpart.alloc(ndata * ob.rot_chunks));
                    SB_TRY(ob.Pbuf.zero(stream));
                    SB_TRY(ob.RA.zero(stream));
                    SB_TRY(ob.RB.zero(stream));
                    SB_TRY(ob.Rres.zero(stream));
                    SB_TRY(ob.Rpart.zero(stream));
                    SB_TRY(raise_smem((const void *)ob.ky.column_fwd, ob.smem_col));
                    SB_TRY(raise_smem((const void *)ob.ky.column_inv, ob.smem_col));
                    SB_TRY(raise_smem((const void *)k_rot_partial<T>, ob.rot_smem));
                }
                if (od.psf_shift) {
                    if (od.kind != 0) return set_err(SB_ERR_ARG, "observation %d: psf_shift needs a ConvolutionRenderer", o);
                    if (od.khat_shared && S > 1) return set_err(SB_ERR_ARG, "observation %d: psf_shift needs one renderer (kernel) per scene", o);
                    if (od.shift_Fx <= 0 || od.shift_Fy <= 0 || (od.shift_Fx & 1)) return set_err(SB_ERR_ARG, "observation %d: bad psf_shift grid %dx%d", o, od.shift_Fy, od.shift_Fx);
                    ob.psf_shift = true, ob.slot0 = psf_slot0[o], ob.shift_Fy = od.shift_Fy, ob.shift_Fx = od.shift_Fx;
                    ob.shift_step = od.shift_step, ob.shift_fixed = od.shift_fixed;
                    SB_TRY(ob.resid.alloc((size_t)S * od.C * desc.Ny * desc.Nx));
                    SB_TRY(ob.resid.zero(stream));
                    if (d_model.n < (size_t)S * C * desc.Ny * desc.Nx) SB_TRY(d_model.alloc((size_t)S * C * desc.Ny * desc.Nx));
                    SB_TRY(d_model.zero(stream));
                }
                SB_TRY(ob.partials.alloc((size_t)S * ob.n_part));
                SB_TRY(ob.partials.zero(stream));
                d.A = nullptr, d.B = ob.G.p, d.Ahat = nullptr, d.khat = ob.khat.p, d.data = ob.data.p, d.weights = ob.weights.p;
                SpecObs<T> &sd = ob.sdev;
                memset(&sd, 0, sizeof sd);
                sd.C = od.C, sd.H = od.H, sd.W = od.W, sd.chan_off = od.chan_off, sd.oy = od.oy, sd.ox = od.ox;
                sd.Fy = Fy, sd.Fx = Fx, sd.Fxc = d.Fxc, sd.Xp = Xp, sd.khat_shared = od.khat_shared;
                sd.X = ob.X.p, sd.khat = ob.khat.p, sd.G = ob.G.p, sd.data = ob.data.p, sd.weights = ob.weights.p;
                sd.tw_x = ob.tw_x.p, sd.tw_y = ob.tw_y.p;
                sd.P = ob.Pbuf.p, sd.T1 = ob.T1buf.p, sd.Ey = ob.Ey.p, sd.Ex = ob.Ex.p, sd.h2 = T(1);
                sd.RA = ob.RA.p, sd.RB = ob.RB.p, sd.Rres = ob.Rres.p, sd.Rpart = ob.Rpart.p, sd.n_chunk = ob.rot_chunks, sd.chunk = ob.rot_chunk;
                sd.cand_start = ob.cand_start.p, sd.cand = ob.cand.p;
                continue;
            }
            d.Kp = d.Fxc, d.Bh = Fy, d.Bw = Fx;
            const size_t ngrid = (size_t)S * od.C * Fy * Fx, ncplx = (size_t)S * od.C * Fy * d.Fxc;
            SB_TRY(ob.A.alloc(ngrid));
            SB_TRY(ob.A.zero(stream));
            ob.n_part = od.C * ((desc.Nx + 31) / 32) * ((desc.Ny + 7) / 8);
            SB_TRY(ob.partials.alloc((size_t)S * ob.n_part));
            SB_TRY(ob.partials.zero(stream));
            if (od.kind == 0) {
                SB_TRY(ob.B.alloc(ngrid));
                SB_TRY(ob.Ahat.alloc(ncplx));
                SB_TRY(ob.khat.alloc(od.khat_shared ? (size_t)od.C * Fy * d.Fxc : ncplx));
                SB_TRY(ob.khat.zero(stream));
                int n[2] = {Fy, Fx};
                SB_CUFFT(cufftPlanMany(&ob.fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, Cx<T>::r2c, S * od.C));
                SB_CUFFT(cufftPlanMany(&ob.inv, 2, n, nullptr, 1, 0, nullptr, 1, 0, Cx<T>::c2r, S * od.C));
                ob.have_plans = true;
                SB_CUFFT(cufftSetStream(ob.fwd, stream));
                SB_CUFFT(cufftSetStream(ob.inv, stream));
                size_t ws = 0;
                cufftGetSize(ob.fwd, &ws);
                dev_bytes += (int64_t)ws;
                cufftGetSize(ob.inv, &ws);
                dev_bytes += (int64_t)ws;
            }
            d.A = ob.A.p, d.B = od.kind == 0 ? ob.B.p : ob.A.p, d.Ahat = ob.Ahat.p, d.khat = ob.khat.p;
            d.data = ob.data.p, d.weights = ob.weights.p;
        }
        SB_TRY(d_stage.alloc(1));
        SB_TRY(reset_counters());
        SB_CUDA(cudaStreamSynchronize(stream));
        dev_bytes += total_bytes();
        return SB_OK;
    }

    int raise_smem(const void *fn, size_t bytes) { return sb::raise_smem_limit(device, fn, bytes); }

    // [R1][R2] table exp(-2 pi i n2 k1 / (R1 R2)) of the two-stage transforms (fft_core.cuh), computed in double
    int upload_twiddles(DevBuf<cplx> &buf, int R1, int R2) {
        const int L = R1 * R2;
        std::vector<cplx> h(L);
        for (int k1 = 0; k1 < R1; ++k1)
            for (int n2 = 0; n2 < R2; ++n2) {
                const sbfft::TwPair t = sbfft::ct_twiddle(n2 * k1, L);
                h[k1 * R2 + n2].x = (T)t.c, h[k1 * R2 + n2].y = (T)(-t.s);
            }
        SB_TRY(buf.alloc(L));
        SB_CUDA(cudaMemcpy(buf.p, h.data(), L * sizeof(cplx), cudaMemcpyHostToDevice));
        return SB_OK;
    }

    // The ExtendedSource chain Monotonicity -> [Symmetry] -> Positivity -> CenterOn -> Normalization("max") (fused_chain_of)
    bool chain_is_fused(int chain, int By, int Bx) const {
        const sb_chain_desc &ch = desc.chains[chain];
        if (ch.repeat != 1 || !(By & 1) || !(Bx & 1)) return false;
        int i = 0;
        if (i >= ch.n_ops || ch.ops[i].code != SB_OP_MONOTONIC) return false;
        ++i;
        if (i < ch.n_ops && ch.ops[i].code == SB_OP_SYMMETRY) ++i;
        if (i >= ch.n_ops || ch.ops[i++].code != SB_OP_POSITIVITY) return false;
        if (i >= ch.n_ops || ch.ops[i++].code != SB_OP_CENTER_ON) return false;
        if (i >= ch.n_ops || ch.ops[i].code != SB_OP_NORMALIZE || ch.ops[i].iarg != 1) return false;
        return i + 1 == ch.n_ops;
    }
    // cap = table entries (DevMono::w_cap); per warp: image, x/psi ring (WARP_RING bytes), spectrum-gradient slots, mbarriers
    static size_t warp_smem_bytes(int G, int npix, int cap) {
        return (size_t)cap * (sizeof(W4<T>) + sizeof(uint2) + sizeof(unsigned short)) + 64 +
               (size_t)G * (SB_FAST_MAXC * sizeof(double) + (size_t)npix * sizeof(T) + WARP_RING + 64);
    }

    // Split the sources between the warp-per-source kernel (k_update_warp), the grouped kernel (k_update_fast) and the generic
    // one (k_update).
    int plan_fast_path() {
        std::map<int, std::vector<int>> by_chain, warp_chain;
        std::vector<int> generic;
        int npix = 0, cap = 0, wnpix = 0, wcap = 0;
        const bool use_warp = getenv("SB_NO_WARP_UPDATE") == nullptr;
        for (int k = 0; k < n_src; ++k) {
            const DevSource &d = h_src[k];
            bool fast = d.kind == 0 && d.chain >= 0 && C <= SB_FAST_MAXC && !d.shifting;
            int tasks = 0, trips = 0;
            if (fast) {
                int n_mono_ops = 0;
                const sb_chain_desc &ch = desc.chains[d.chain];
                for (int i = 0; i < ch.n_ops; ++i)
                    if (ch.ops[i].code == SB_OP_MONOTONIC) {
                        const HostMono &h = hmonos[ch.ops[i].iarg];
                        ++n_mono_ops;
                        tasks = h.n_tasks;
                        for (int L = 0; L < h.n_levels; ++L) trips += (h.level_start[L + 1] - h.level_start[L] + 31) / 32;
                        if (h.nb != 4 || h.n_levels + 1 > 512) fast = false;
                    }
                if (n_mono_ops > 1) fast = false;
            }
            if (fast) {
                // must fit next to one image even with a single group per CTA
                const size_t need = fast_smem_bytes(1, (d.By * d.Bx + 4) & ~3, (tasks + 7) & ~7);
                if (need > 200 * 1024) fast = false;
                // the shared-memory table addresses pixels by 16-bit byte offsets (spare cell included)
                if ((size_t)(d.By * d.Bx + 1) * sizeof(T) > 65535) fast = false;
            }
            const bool warp = fast && use_warp && chain_is_fused(d.chain, d.By, d.Bx) && d.By * d.Bx <= 32 * WARP_NPT && trips <= 1020 &&
                              warp_smem_bytes(4, (d.By * d.Bx + 4) & ~3, 32 * (trips + (trips & 1) + 1)) <= 200 * 1024;
            if (warp) {
                warp_chain[d.chain].push_back(k);
                wnpix = std::max(wnpix, (d.By * d.Bx + 4) & ~3);
                wcap = std::max(wcap, 32 * (trips + (trips & 1) + 1));
            } else if (fast) {
                by_chain[d.chain].push_back(k);
                npix = std::max(npix, (d.By * d.Bx + 4) & ~3); // + the spare zero cell of group_sweep
                cap = std::max(cap, (tasks + 7) & ~7);
            } else
                generic.push_back(k);
        }
        fast_npix = npix, fast_table_cap = std::max(cap, 8);
        warp_npix = wnpix, warp_cap = std::max(wcap, 8);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        // Every SM gets the same number of groups (= sources): the kernels are bound by the latency of the per-source
        // proximal loop, so balance matters more than occupancy.  One CTA per SM.
        auto pick_G = [&](size_t n_items, int gmax, auto smem_of) {
            if (!n_items) return 0;
            const size_t slots = (size_t)sms;
            const int waves = (int)((n_items + slots * gmax - 1) / (slots * gmax));
            int G = (int)((n_items + slots * waves - 1) / (slots * waves));
            G = std::max(1, std::min(gmax, G));
            while (G > 1 && smem_of(G) > (size_t)220 * 1024) --G;
            return smem_of(G) <= (size_t)220 * 1024 ? G : 0;
        };
        auto make_groups = [](const std::map<int, std::vector<int>> &chains, int G, std::vector<int> &groups) {
            for (auto &kv : chains) {
                const std::vector<int> &v = kv.second;
                for (size_t i = 0; i < v.size(); i += G)
                    for (int j = 0; j < G; ++j) groups.push_back(i + j < v.size() ? v[i + j] : -1);
            }
        };
        // ---- warp-per-source kernel: up to WARP_MAXT / 32 warps per CTA (register-resident iterate)
        {
            size_t n_warp = 0;
            for (auto &kv : warp_chain) n_warp += kv.second.size();
            // Persistent warps: one CTA per SM (fewer when the batch is small), G warps each; the CTAs are dealt to the chains in
            // proportion to their source counts and the warps of a chain's CTAs claim its sources from one counter.
            {
                const int gmax = WARP_MAXT / 32;
                int G = n_warp ? (int)std::min<size_t>(gmax, (n_warp + sms - 1) / sms) : 0;
                while (G > 1 && warp_smem_bytes(G, warp_npix, warp_cap) > (size_t)220 * 1024) --G;
                if (G && warp_smem_bytes(G, warp_npix, warp_cap) > (size_t)220 * 1024) G = 0;
                warp_G = G;
            }
            std::vector<int> groups;
            std::vector<int4> segs;
            n_warp_chains = 0;
            if (warp_G) {
                const size_t n_cta_target = std::min<size_t>(sms, (n_warp + warp_G - 1) / warp_G);
                for (auto &kv : warp_chain) {
                    const std::vector<int> &v = kv.second;
                    const size_t ctas = std::max<size_t>(1, std::min<size_t>((v.size() + warp_G - 1) / warp_G,
                                                                             (n_cta_target * v.size() + n_warp / 2) / n_warp));
                    for (size_t c = 0; c < ctas; ++c) segs.push_back(make_int4((int)groups.size(), (int)v.size(), n_warp_chains, 0));
                    groups.insert(groups.end(), v.begin(), v.end());
                    ++n_warp_chains;
                }
            } else
                for (auto &kv : warp_chain) by_chain[kv.first].insert(by_chain[kv.first].end(), kv.second.begin(), kv.second.end());
            n_warp_cta = (int)segs.size();
            warp_smem = warp_G ? warp_smem_bytes(warp_G, warp_npix, warp_cap) : 0;
            SB_TRY(d_warp_groups.alloc(std::max<size_t>(groups.size(), 1)));
            SB_TRY(d_warp_segs.alloc(std::max<size_t>(segs.size(), 1)));
            SB_TRY(d_warp_ctr.alloc(std::max(n_warp_chains, 1)));
            if (!groups.empty()) SB_CUDA(cudaMemcpy(d_warp_groups.p, groups.data(), groups.size() * sizeof(int), cudaMemcpyHostToDevice));
            if (!segs.empty()) SB_CUDA(cudaMemcpy(d_warp_segs.p, segs.data(), segs.size() * sizeof(int4), cudaMemcpyHostToDevice));
            if (n_warp_cta) {
                SB_TRY(d_xp.alloc((size_t)n_warp_cta * warp_G * (32 * WARP_NPT))); // one padded, 16-byte aligned slot per warp
                SB_TRY(raise_smem((const void *)k_update_warp<T, WARP_NPT, WARP_MAXT>, warp_smem));
            }
        }
        // ---- grouped kernel: up to 16 groups of GT = 64 threads (default), or 8 groups of 128 (SB_UPDATE_GROUP=128)
        fast_G = 0;
        {
            const char *gt = getenv("SB_UPDATE_GROUP");
            fast_GT = (gt && atoi(gt) == 128) ? 128 : (gt && atoi(gt) == 32) ? 32 : 64;
            size_t n_fast = 0;
            for (auto &kv : by_chain) n_fast += kv.second.size();
            fast_G = pick_G(n_fast, 1024 / fast_GT, [&](int G) { return fast_smem_bytes(G, fast_npix, fast_table_cap); });
        }
        std::vector<int> groups;
        if (fast_G) {
            make_groups(by_chain, fast_G, groups);
        } else {
            for (auto &kv : by_chain) generic.insert(generic.end(), kv.second.begin(), kv.second.end());
            std::sort(generic.begin(), generic.end());
        }
        n_fast_cta = fast_G ? (int)(groups.size() / fast_G) : 0;
        n_generic = (int)generic.size();
        fast_smem = fast_G ? fast_smem_bytes(fast_G, fast_npix, fast_table_cap) : 0;
        SB_TRY(d_work.alloc(std::max<size_t>(generic.size(), 1)));
        SB_TRY(d_fast_groups.alloc(std::max<size_t>(groups.size(), 1)));
        if (!generic.empty()) SB_CUDA(cudaMemcpy(d_work.p, generic.data(), generic.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (!groups.empty()) SB_CUDA(cudaMemcpy(d_fast_groups.p, groups.data(), groups.size() * sizeof(int), cudaMemcpyHostToDevice));
        SB_TRY(d_scratch_x.alloc(std::max<long long>(n_morph, 1))); // both update kernels stream x and psi through these
        SB_TRY(d_scratch_ps.alloc(std::max<long long>(n_morph, 1)));
        if (n_fast_cta) {
            SB_TRY(raise_smem(fast_GT == 128 ? (const void *)k_update_fast<T, 128> : (const void *)k_update_fast<T, 64>, fast_smem));
            SB_TRY(raise_smem((const void *)k_update_fast<T, 32, 448>, fast_smem));
            SB_TRY(raise_smem((const void *)k_update_fast<T, 32>, fast_smem));
            SB_TRY(raise_smem((const void *)k_update_fast<T, 64, 832>, fast_smem));
        }
        return SB_OK;
    }
    static size_t fast_smem_bytes(int G, int npix, int cap) {
        return (size_t)cap * (sizeof(W4<T>) + sizeof(uint2)) + (size_t)G * (16 + SB_MAXC) * sizeof(double) + 512 * sizeof(int) +
               (size_t)((cap + 7) & ~7) * sizeof(unsigned short) + (size_t)G * npix * sizeof(T);
    }

    size_t update_smem() const { return ((size_t)npix_max + 3 * (size_t)npix_shift) * sizeof(T) + (40 + SB_MAXC) * sizeof(double); }

    int ensure_loss_cap(int cap) {
        if (cap <= loss_cap) return SB_OK;
        SB_CUDA(cudaStreamSynchronize(stream));
        SB_TRY(d_loss.alloc((size_t)S * cap));
        SB_TRY(d_loss.zero(stream));
        loss_cap = cap;
        have_graph = false; // cap is baked into kernel arguments
        return SB_OK;
    }

    // the optional per-scene tables (sb_plan_scene_control) travel in the kernel arguments: plain fits run without them
    void drop_scene_tables() {
        if (use_limit || use_prox_iter) have_graph = false;
        use_limit = use_prox_iter = false;
    }

    int reset_counters() {
        SB_TRY(d_done.zero(stream));
        SB_TRY(d_niter.zero(stream));
        SB_TRY(d_status.zero(stream));
        SB_TRY(d_it.zero(stream));
        SB_TRY(d_state.zero(stream));
        SB_TRY(d_nactive_next.zero(stream));
        SB_CUDA(cudaMemcpyAsync(d_nactive.p, &S, sizeof(int), cudaMemcpyHostToDevice, stream));
        return SB_OK;
    }

    // ---------------------------------------------------------------------------------------------
    int upload_observation(int o, const void *data, const void *weights, int elem_bytes, const double *khat, const double *loss_const) override {
        if (o < 0 || o >= (int)obs.size()) return set_err(SB_ERR_ARG, "observation index %d out of range", o);
        SB_CUDA(cudaSetDevice(device));
        Obs &ob = *obs[o];
        const size_t nd = ob.data.n;
        const bool same = (size_t)elem_bytes == sizeof(T);
        // staging for the cast on the device (kept between calls, see Obs::stage_f)
        if (!same && (data || weights)) {
            if (elem_bytes == 4 && ob.stage_f.n < nd) SB_TRY(ob.stage_f.alloc(nd));
            if (elem_bytes == 8 && ob.stage_z.n < (nd + 1) / 2) SB_TRY(ob.stage_z.alloc((nd + 1) / 2));
        }
        const void *srcs[2] = {data, weights};
        T *dsts[2] = {ob.data.p, ob.weights.p};
        for (int i = 0; i < 2; ++i) {
            if (!srcs[i]) continue;
            if (same) {
                SB_CUDA(cudaMemcpyAsync(dsts[i], srcs[i], nd * sizeof(T), cudaMemcpyHostToDevice, stream));
            } else if (elem_bytes == 4) {
                SB_CUDA(cudaMemcpyAsync(ob.stage_f.p, srcs[i], nd * sizeof(float), cudaMemcpyHostToDevice, stream));
                k_cast<float, T><<<grid_for(nd), 256, 0, stream>>>(ob.stage_f.p, dsts[i], (long long)nd);
                SB_CUDA(cudaGetLastError());
            } else {
                double *st = reinterpret_cast<double *>(ob.stage_z.p);
                SB_CUDA(cudaMemcpyAsync(st, srcs[i], nd * sizeof(double), cudaMemcpyHostToDevice, stream));
                k_cast<double, T><<<grid_for(nd), 256, 0, stream>>>(st, dsts[i], (long long)nd);
                SB_CUDA(cudaGetLastError());
            }
        }
        if (khat && (ob.dev.kind == 0 || ob.dev.kind == 2 || ob.dev.kind == 3)) {
            DevBuf<double2> &ks = ob.stage_z;
            const size_t nk = (size_t)(ob.dev.khat_shared ? 1 : S) * ob.dev.C * ob.dev.Fy * ob.dev.Fxc;
            if (ks.n < nk) SB_TRY(ks.alloc(nk));
            SB_CUDA(cudaMemcpyAsync(ks.p, khat, nk * sizeof(double2), cudaMemcpyHostToDevice, stream));
            const long long nrows = (long long)(ob.dev.khat_shared ? 1 : S) * ob.dev.C * ob.dev.Fy;
            k_cast_scale_cplx_pitched<T><<<grid_for(nrows * ob.dev.Fxc), 256, 0, stream>>>(ks.p, ob.khat.p, nrows, ob.dev.Fxc, ob.dev.Kp,
                                                                                         1.0 / ((double)ob.dev.Fy * ob.dev.Fx));
            SB_CUDA(cudaGetLastError());
            SB_CUDA(cudaStreamSynchronize(stream));
        }
        if (loss_const) {
            ob.loss_const.assign(loss_const, loss_const + S);
            std::vector<double> tot(S, 0.0);
            for (auto &q : obs)
                for (int s = 0; s < S; ++s) tot[s] += q->loss_const[s];
            SB_CUDA(cudaMemcpyAsync(d_loss_const.p, tot.data(), S * sizeof(double), cudaMemcpyHostToDevice, stream));
        }
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }

    int upload_kernels(int o, const double *ker, int Py, int Px, int y0, int x0) override {
        if (o < 0 || o >= (int)obs.size()) return set_err(SB_ERR_ARG, "observation index %d out of range", o);
        Obs &ob = *obs[o];
        const DevObs<T> &d = ob.dev;
        if (d.kind != 0) return set_err(SB_ERR_ARG, "observation %d takes no kernel image (NullRenderer, or a resampling observation whose K^ is uploaded as is)", o);
        if (!ker || Py <= 0 || Px <= 0 || Py > d.Fy || Px > d.Fx) return set_err(SB_ERR_ARG, "bad kernel image %dx%d", Py, Px);
        if (d.Fy < desc.Ny + std::max(-y0, Py - 1 + y0) || d.Fx < desc.Nx + std::max(-x0, Px - 1 + x0) || y0 > 0 || x0 > 0 ||
            Py - 1 + y0 < 0 || Px - 1 + x0 < 0)
            return set_err(SB_ERR_ARG, "FFT grid %dx%d too small for frame %dx%d with a %dx%d kernel at (%d,%d)", d.Fy, d.Fx,
                           desc.Ny, desc.Nx, Py, Px, y0, x0);
        SB_CUDA(cudaSetDevice(device));
        const int nk = d.khat_shared ? 1 : S;
        const size_t per_img = (size_t)Py * Px, per_grid = (size_t)d.Fy * d.Fx, per_spec = (size_t)d.Fy * d.Fxc;
        if (!ob.have_kplan) {
            ob.kchunk = std::min(nk, 8);
            int n[2] = {d.Fy, d.Fx};
            SB_CUFFT(cufftPlanMany(&ob.kplan, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, ob.kchunk * d.C));
            ob.have_kplan = true;
            SB_CUFFT(cufftSetStream(ob.kplan, stream));
            SB_TRY(ob.kgrid.alloc((size_t)ob.kchunk * d.C * per_grid));
            SB_TRY(ob.kspec.alloc((size_t)ob.kchunk * d.C * per_spec));
        }
        if (ob.kimg.n < (size_t)nk * d.C * per_img) SB_TRY(ob.kimg.alloc((size_t)nk * d.C * per_img));
        SB_CUDA(cudaMemcpyAsync(ob.kimg.p, ker, (size_t)nk * d.C * per_img * sizeof(double), cudaMemcpyHostToDevice, stream));
        ob.Py = Py, ob.Px = Px, ob.ky0 = y0, ob.kx0 = x0, ob.have_kernel = true;
        if (ob.psf_shift) { // K^ follows the fitted offset: shifted kernel (and its derivatives) for the current parameter
            const size_t nimg = (size_t)S * d.C * per_img;
            if (ob.ks.n < nimg) {
                SB_TRY(ob.ks.alloc(nimg));
                SB_TRY(ob.kd0.alloc(nimg));
                SB_TRY(ob.kd1.alloc(nimg));
                SB_TRY(ob.gk.alloc(nimg));
            }
            const size_t smem = (size_t)(8 * (Py + Px) + 3 * Py * Px) * sizeof(double);
            if (smem > 200 * 1024) return set_err(SB_ERR_ARG, "psf_shift: kernel image %dx%d too large", Py, Px);
            SB_TRY(raise_smem((const void *)k_psf_kernel<T>, smem));
            SB_TRY(refresh_psf(o));
        } else
            SB_TRY(khat_from_images(ob, ob.kimg.p, nk));
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }

    // kernel images [n][C][Py][Px] (device, double) -> K^/(Fy Fx) in the plan's precision (pad, wrap, transform in double)
    int khat_from_images(Obs &ob, const double *img, int nk) {
        const DevObs<T> &d = ob.dev;
        const size_t per_img = (size_t)ob.Py * ob.Px, per_spec = (size_t)d.Fy * d.Fxc;
        for (int s0 = 0; s0 < nk; s0 += ob.kchunk) {
            const int nb = std::min(ob.kchunk, nk - s0);
            SB_TRY(ob.kgrid.zero(stream));
            const long long nimg = (long long)nb * d.C * per_img;
            k_embed_kernel<<<grid_for(nimg), 256, 0, stream>>>(img + (size_t)s0 * d.C * per_img, ob.kgrid.p, nb * d.C, ob.Py, ob.Px, d.Fy, d.Fx,
                                                             ob.ky0, ob.kx0);
            SB_CUDA(cudaGetLastError());
            SB_CUFFT(cufftExecD2Z(ob.kplan, ob.kgrid.p, ob.kspec.p));
            const long long nspec = (long long)nb * d.C * per_spec, nrows = (long long)nb * d.C * d.Fy;
            k_cast_scale_cplx_pitched<T><<<grid_for(nspec), 256, 0, stream>>>(ob.kspec.p, ob.khat.p + (size_t)s0 * d.C * d.Fy * d.Kp, nrows,
                                                                            d.Fxc, d.Kp, 1.0 / ((double)d.Fy * d.Fx));
            SB_CUDA(cudaGetLastError());
        }
        return SB_OK;
    }

    PsfShiftArgs<T> psf_args(int o, int mode) {
        Obs &ob = *obs[o];
        PsfShiftArgs<T> a;
        memset(&a, 0, sizeof a);
        a.S = S, a.C = ob.dev.C, a.Py = ob.Py, a.Px = ob.Px, a.Fy = ob.shift_Fy, a.Fx = ob.shift_Fx;
        a.Ny = desc.Ny, a.Nx = desc.Nx, a.chan_off = ob.dev.chan_off, a.Cm = C, a.slot0 = ob.slot0;
        a.kernel_shared = ob.dev.khat_shared, a.fixed = ob.shift_fixed, a.mode = mode, a.step = ob.shift_step;
        a.kimg = ob.kimg.p, a.ks = ob.ks.p, a.kd0 = ob.kd0.p, a.kd1 = ob.kd1.p, a.gk = ob.gk.p;
        a.model = d_model.p, a.resid = ob.resid.p;
        a.center = d_center.p, a.cen_m = d_cen_m.p, a.cen_v = d_cen_v.p, a.cen_vhat = d_cen_vhat.p, a.g_center = d_gcenter.p;
        a.it_ptr = d_it.p, a.done = d_done.p, a.status = d_status.p, a.fs = cur_fs;
        return a;
    }
    // shifted kernel for the current offsets -> K^ (every scene has its own kernel: psf_shift plans are never khat_shared for S > 1)
    int refresh_psf(int o) {
        Obs &ob = *obs[o];
        if (!ob.have_kernel) return SB_OK;
        PsfShiftArgs<T> pa = psf_args(o, 0);
        const size_t smem = (size_t)(8 * (ob.Py + ob.Px) + 3 * ob.Py * ob.Px) * sizeof(double);
        k_psf_kernel<T><<<S * ob.dev.C, 128, smem, stream>>>(pa);
        SB_CUDA(cudaGetLastError());
        SB_TRY(khat_from_images(ob, ob.ks.p, S));
        return SB_OK;
    }

    int upload_resampling(int o, const double *ey, const double *ex, double h2) override {
        if (o < 0 || o >= (int)obs.size()) return set_err(SB_ERR_ARG, "observation index %d out of range", o);
        Obs &ob = *obs[o];
        if (ob.dev.kind != 2 || !ey || !ex) return set_err(SB_ERR_ARG, "observation %d is not a resampling observation", o);
        SB_CUDA(cudaSetDevice(device));
        DevBuf<double2> &st = ob.stage_z;
        if (st.n < std::max(ob.Ey.n, ob.Ex.n)) SB_TRY(st.alloc(std::max(ob.Ey.n, ob.Ex.n)));
        SB_CUDA(cudaMemcpyAsync(st.p, ey, ob.Ey.n * sizeof(double2), cudaMemcpyHostToDevice, stream));
        k_cast_scale_cplx<T><<<grid_for(ob.Ey.n), 256, 0, stream>>>(st.p, ob.Ey.p, (long long)ob.Ey.n, 1.0);
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaStreamSynchronize(stream));
        SB_CUDA(cudaMemcpyAsync(st.p, ex, ob.Ex.n * sizeof(double2), cudaMemcpyHostToDevice, stream));
        k_cast_scale_cplx<T><<<grid_for(ob.Ex.n), 256, 0, stream>>>(st.p, ob.Ex.p, (long long)ob.Ex.n, 1.0);
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaStreamSynchronize(stream));
        if (ob.sdev.h2 != (T)h2) have_graph = false; // h2 travels in the kernel arguments
        ob.sdev.h2 = (T)h2;
        return SB_OK;
    }

    // rotated resampling: A [H][Fy][Fx/2+1], B [W][Fy][Fx/2+1] complex128 -> pitched tables in the plan's precision
    int upload_resampling_rot(int o, const double *ra, const double *rb, double h2) override {
        if (o < 0 || o >= (int)obs.size()) return set_err(SB_ERR_ARG, "observation index %d out of range", o);
        Obs &ob = *obs[o];
        if (ob.dev.kind != 3 || !ra || !rb) return set_err(SB_ERR_ARG, "observation %d is not a rotated resampling observation", o);
        SB_CUDA(cudaSetDevice(device));
        const DevObs<T> &d = ob.dev;
        const size_t per = (size_t)d.Fy * d.Fxc, na = (size_t)d.H * per, nb = (size_t)d.W * per;
        DevBuf<double2> &st = ob.stage_z;
        if (st.n < std::max(na, nb)) SB_TRY(st.alloc(std::max(na, nb)));
        const double *src[2] = {ra, rb};
        const size_t cnt[2] = {na, nb};
        cplx *dst[2] = {ob.RA.p, ob.RB.p};
        for (int q = 0; q < 2; ++q) {
            SB_CUDA(cudaMemcpyAsync(st.p, src[q], cnt[q] * sizeof(double2), cudaMemcpyHostToDevice, stream));
            k_cast_scale_cplx_pitched<T><<<grid_for(cnt[q]), 256, 0, stream>>>(st.p, dst[q], (long long)(cnt[q] / d.Fxc), d.Fxc, d.Kp, 1.0);
            SB_CUDA(cudaGetLastError());
            SB_CUDA(cudaStreamSynchronize(stream));
        }
        if (ob.sdev.h2 != (T)h2) have_graph = false; // h2 travels in the kernel arguments
        ob.sdev.h2 = (T)h2;
        return SB_OK;
    }

    int zero_state() override {
        SB_CUDA(cudaSetDevice(device));
        DevBuf<double> *dz[] = {&d_sed_m, &d_sed_v, &d_sed_vhat, &d_cen_m, &d_cen_v, &d_cen_vhat};
        for (auto *bf : dz) SB_TRY(bf->zero(stream));
        DevBuf<T> *dm[] = {&d_morph_m, &d_morph_v, &d_morph_vhat};
        for (auto *bf : dm) SB_TRY(bf->zero(stream));
        return SB_OK;
    }

    DevBuf<double> stage_d;
    int upload_T(T *dst, const double *src, size_t n) {
        if (!src || n == 0) return SB_OK;
        if (sizeof(T) == 8) {
            SB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        } else {
            if (stage_d.n < n) SB_TRY(stage_d.alloc(n));
            SB_CUDA(cudaMemcpyAsync(stage_d.p, src, n * sizeof(double), cudaMemcpyHostToDevice, stream));
            k_cast<double, T><<<grid_for(n), 256, 0, stream>>>(stage_d.p, dst, (long long)n);
            SB_CUDA(cudaGetLastError());
            SB_CUDA(cudaStreamSynchronize(stream));
        }
        return SB_OK;
    }
    int download_T(double *dst, const T *src, size_t n) {
        if (!dst || n == 0) return SB_OK;
        if (sizeof(T) == 8) {
            SB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        } else {
            if (stage_d.n < n) SB_TRY(stage_d.alloc(n));
            k_cast<T, double><<<grid_for(n), 256, 0, stream>>>(src, stage_d.p, (long long)n);
            SB_CUDA(cudaGetLastError());
            SB_CUDA(cudaMemcpyAsync(dst, stage_d.p, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        }
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }

    int upload_params(int which, const double *sed, const double *morph, const double *center) override {
        if (which < 0 || which > 3) return set_err(SB_ERR_ARG, "which=%d", which);
        SB_CUDA(cudaSetDevice(device));
        double *ds[4] = {d_sed.p, d_sed_m.p, d_sed_v.p, d_sed_vhat.p};
        T *dm[4] = {d_morph.p, d_morph_m.p, d_morph_v.p, d_morph_vhat.p};
        double *dc[4] = {d_center.p, d_cen_m.p, d_cen_v.p, d_cen_vhat.p};
        if (sed && n_src) SB_CUDA(cudaMemcpyAsync(ds[which], sed, (size_t)n_src * C * sizeof(double), cudaMemcpyHostToDevice, stream));
        SB_TRY(upload_T(dm[which], morph, (size_t)n_morph));
        if (center && n_point) SB_CUDA(cudaMemcpyAsync(dc[which], center, (size_t)n_point * 2 * sizeof(double), cudaMemcpyHostToDevice, stream));
        if (which == 0 && center && n_point) {
            UpdateArgs<T> ua = update_args(0);
            k_point_morph<T><<<n_src, 128, 0, stream>>>(ua, n_src);
            SB_CUDA(cudaGetLastError());
        }
        if (which == 0 && n_shift) SB_TRY(launch_shift_apply());
        if (which == 0 && center)
            for (size_t o = 0; o < obs.size(); ++o)
                if (obs[o]->psf_shift) SB_TRY(refresh_psf((int)o));
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }
    int launch_shift_apply() {
        UpdateArgs<T> ua = update_args(0);
        k_shift_apply<T><<<n_shift, 128, (size_t)3 * npix_shift * sizeof(T), stream>>>(ua);
        SB_CUDA(cudaGetLastError());
        return SB_OK;
    }
    int download_params(int which, double *sed, double *morph, double *center) override {
        if (which < 0 || which > 3) return set_err(SB_ERR_ARG, "which=%d", which);
        SB_CUDA(cudaSetDevice(device));
        double *ds[4] = {d_sed.p, d_sed_m.p, d_sed_v.p, d_sed_vhat.p};
        T *dm[4] = {d_morph.p, d_morph_m.p, d_morph_v.p, d_morph_vhat.p};
        double *dc[4] = {d_center.p, d_cen_m.p, d_cen_v.p, d_cen_vhat.p};
        if (sed && n_src) SB_CUDA(cudaMemcpyAsync(sed, ds[which], (size_t)n_src * C * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (center && n_point) SB_CUDA(cudaMemcpyAsync(center, dc[which], (size_t)n_point * 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        SB_TRY(download_T(morph, dm[which], (size_t)n_morph));
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }
    int spectral_mode() const override { return fused ? 1 : 0; }
    // diagnostic: how many of the <= prox_max_iter proximal sub-iterations the grouped update kernel actually runs
    DevBuf<unsigned long long> d_prox_hist;
    int prox_histogram(int enable, int64_t *out16) override {
        SB_CUDA(cudaSetDevice(device));
        if (enable && !d_prox_hist.p) {
            SB_TRY(d_prox_hist.alloc(16));
            have_graph = false; // the pointer travels in the kernel arguments
        }
        if (d_prox_hist.p && out16) {
            SB_CUDA(cudaMemcpyAsync(out16, d_prox_hist.p, 16 * sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
            SB_CUDA(cudaStreamSynchronize(stream));
        }
        if (d_prox_hist.p && enable) SB_TRY(d_prox_hist.zero(stream));
        if (!enable && d_prox_hist.p) {
            SB_CUDA(cudaStreamSynchronize(stream));
            d_prox_hist.release();
            have_graph = false;
        }
        return SB_OK;
    }
    int device_params(void **sed, int64_t *n_sed, void **morph, int64_t *nm, int *elem_bytes) override {
        if (sed) *sed = d_sed.p;
        if (n_sed) *n_sed = (int64_t)n_src * C;
        if (morph) *morph = d_morph.p;
        if (nm) *nm = n_morph;
        if (elem_bytes) *elem_bytes = (int)sizeof(T);
        return SB_OK;
    }

    // ---------------------------------------------------------------------------------------------
    static FitScalars scalars(const sb_fit_opts *o) {
        FitScalars f;
        memset(&f, 0, sizeof f);
        f.prox_max_iter = o->prox_max_iter, f.min_iter = o->min_iter, f.fixed_iterations = o->fixed_iterations;
        f.overwrite_vhat_at_it0 = o->overwrite_vhat_at_it0;
        f.pause_every = o->pause_every;
        f.e_rel = o->e_rel, f.b1 = o->b1, f.b2 = o->b2, f.eps = o->eps;
        return f;
    }
    FitScalars cur_fs;

    UpdateArgs<T> update_args(int mode) {
        UpdateArgs<T> a;
        memset(&a, 0, sizeof a);
        a.src = d_src.p, a.C = C, a.Ny = desc.Ny, a.Nx = desc.Nx, a.n_obs = (int)obs.size();
        for (size_t o = 0; o < obs.size(); ++o) a.obs[o] = obs[o]->dev;
        a.sed = d_sed.p, a.sed_m = d_sed_m.p, a.sed_v = d_sed_v.p, a.sed_vhat = d_sed_vhat.p;
        a.morph = d_morph.p, a.morph_m = d_morph_m.p, a.morph_v = d_morph_v.p, a.morph_vhat = d_morph_vhat.p;
        a.center = d_center.p, a.cen_m = d_cen_m.p, a.cen_v = d_cen_v.p, a.cen_vhat = d_cen_vhat.p;
        a.pmorph = d_pmorph.p, a.chains = d_chains.p, a.monos = d_monos.p;
        a.it_ptr = d_it.p, a.done = d_done.p, a.status = d_status.p;
        a.prox_iter = use_prox_iter ? d_prox_iter.p : nullptr;
        a.fs = cur_fs;
        a.psf_b = desc.psf_boxsize;
        memcpy(a.psf_sigma, desc.psf_sigma, sizeof a.psf_sigma);
        a.npix_max = npix_max, a.npix_shift = npix_shift, a.mode = mode;
        a.g_sed = d_gsed.p, a.g_morph = d_gmorph.p, a.g_center = d_gcenter.p;
        a.work = nullptr, a.fast_groups = d_fast_groups.p, a.fast_G = fast_G, a.fast_npix = fast_npix, a.fast_table_cap = fast_table_cap;
        a.scratch_x = d_scratch_x.p, a.scratch_ps = d_scratch_ps.p;
        a.prox_hist = d_prox_hist.p;
        a.smorph = d_smorph.p, a.toep = d_toep.p, a.toep_len = toep_len, a.shift_list = d_shift_list.p;
        return a;
    }

    struct StageTimer {
        std::vector<cudaEvent_t> ev;
        bool on = false;
    };

    // Enqueue one iteration.  mode 0 = full iteration, 1 = forward + gradients only (evaluate).
    int enqueue_iteration(int mode, T *model_out, T *rendered_out, int rendered_obs, std::vector<cudaEvent_t> *marks) {
        auto mark = [&](void) {
            if (marks) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                cudaEventRecord(e, stream);
                marks->push_back(e);
            }
        };
        int nk = 0, nf = 0;
        mark();
        if (!fused) {
            RenderArgs<T> ra;
            memset(&ra, 0, sizeof ra);
            ra.src = d_src.p, ra.scene_src_start = d_start.p, ra.sed = d_sed.p, ra.morph = d_morph.p, ra.pmorph = d_pmorph.p;
            ra.smorph = d_smorph.p;
            ra.C = C, ra.Ny = desc.Ny, ra.Nx = desc.Nx, ra.n_obs = (int)obs.size();
            for (size_t o = 0; o < obs.size(); ++o) ra.obs[o] = obs[o]->dev;
            ra.done = d_done.p;
            ra.model_out = model_out;
            dim3 grid((desc.Nx + 31) / 32, (desc.Ny + 7) / 8, S), block(32, 8);
            k_render<T><<<grid, block, 0, stream>>>(ra);
            SB_CUDA(cudaGetLastError());
            ++nk;
        }
        mark();
        // fused spectral path: marks after {render+rowFFT, column, -, rowIFFT+residual+rowFFT, -, column*, rowIFFT}
        for (size_t o = 0; fused && o < obs.size(); ++o) {
            Obs &ob = *obs[o];
            SpecArgs<T> sa;
            memset(&sa, 0, sizeof sa);
            sa.ob = ob.sdev, sa.Ny = desc.Ny, sa.Nx = desc.Nx, sa.Cm = C, sa.npair = ob.npair, sa.cb = ob.cb, sa.done = d_done.p;
            sa.src = d_src.p, sa.scene_src_start = d_start.p, sa.sed = d_sed.p, sa.morph = d_morph.p, sa.pmorph = d_pmorph.p;
            sa.smorph = d_smorph.p;
            sa.model_out = ob.psf_shift ? d_model.p : model_out, sa.partials = ob.partials.p;
            sa.resid_out = ob.psf_shift ? ob.resid.p : nullptr;
            sa.magic_nx = 0xffffffffu / (unsigned)desc.Nx + 1u;
            sa.rendered_out = ((int)o == rendered_obs) ? rendered_out : nullptr;
            const dim3 rgrid((desc.Ny + 2 * ob.npair - 1) / (2 * ob.npair), S, (ob.sdev.C + ob.cb - 1) / ob.cb);
            const dim3 cgrid((ob.sdev.Fxc + ob.ky.NBcol - 1) / ob.ky.NBcol, S * ob.sdev.C);
            const int cthreads = ob.ky.NBcol * std::max(ob.ky.R1, ob.ky.R2);
            (ob.render_two ? ob.kx.render2 : ob.kx.render)<<<rgrid, ob.row_threads, ob.smem_render, stream>>>(sa);
            SB_CUDA(cudaGetLastError());
            mark();
            if (ob.dev.kind == 3) { // rotated resampling: M^ -> K^ conj(M^) -> contraction with A_i B_j -> residual -> K^ sum R A B -> G
                const SpecObs<T> &sd = ob.sdev;
                const int K = sd.Fy * sd.Xp;
                ob.ky.column_fwd<<<cgrid, cthreads, ob.smem_col, stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark();
                k_rot_partial<T><<<dim3(sd.n_chunk, (sd.C + ob.rot_cb - 1) / ob.rot_cb, S), 32 * ob.rot_cb, ob.rot_smem, stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark();
                k_rot_residual<T><<<S * sd.C, 256, 0, stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark();
                k_rot_adjoint<T><<<dim3((K + 127) / 128, S * sd.C), 128, (size_t)sd.H * 32 * sizeof(T), stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark();
                ob.ky.column_inv<<<cgrid, cthreads, ob.smem_col, stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark();
                ob.kx.grad<<<rgrid, ob.row_threads, ob.smem_row, stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark();
                nk += 7;
                continue;
            }
            if (ob.dev.kind == 2) { // resampling observation: M^ -> K^ conj(M^) -> Ey . -> LR, residual -> . Ex, Ey^T . -> K^ Q^ -> G
                const SpecObs<T> &sd = ob.sdev;
                const dim3 tgrid((sd.Fxc + 127) / 128, (sd.H + SB_RS_ROWS - 1) / SB_RS_ROWS, S * sd.C);
                const dim3 qgrid((sd.Fxc + 127) / 128, (sd.Fy + SB_RS_KY - 1) / SB_RS_KY, S * sd.C);
                const size_t t1_smem = (size_t)sd.Fy * SB_RS_ROWS * sizeof(cplx), q_smem = (size_t)sd.H * SB_RS_KY * sizeof(cplx);
                ob.ky.column_fwd<<<cgrid, cthreads, ob.smem_col, stream>>>(sa);
                k_resample_t1<T><<<tgrid, 128, t1_smem, stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark(), mark();
                k_resample_lr<T><<<S * sd.C, 256, resample_lr_smem(sd.H, sd.W, sd.Fxc), stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark(), mark();
                k_resample_q<T><<<qgrid, 128, q_smem, stream>>>(sa);
                ob.ky.column_inv<<<cgrid, cthreads, ob.smem_col, stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark();
                ob.kx.grad<<<rgrid, ob.row_threads, ob.smem_row, stream>>>(sa);
                SB_CUDA(cudaGetLastError());
                mark();
                nk += 7;
                continue;
            }
            sa.conj = 0;
            if (ob.use_tma)
                ob.ky.column_tma<<<cgrid, cthreads, ob.smem_col_tma, stream>>>(sa, ob.tmX, ob.tmK);
            else
                ob.ky.column<<<cgrid, cthreads, ob.smem_col, stream>>>(sa);
            SB_CUDA(cudaGetLastError());
            mark(), mark();
            (ob.psf_shift ? ob.kx.residual_r : ob.kx.residual)<<<rgrid, ob.row_threads, ob.smem_row, stream>>>(sa);
            SB_CUDA(cudaGetLastError());
            mark(), mark();
            sa.conj = 1;
            if (ob.use_tma)
                ob.ky.column_tma<<<cgrid, cthreads, ob.smem_col_tma, stream>>>(sa, ob.tmX, ob.tmK);
            else
                ob.ky.column<<<cgrid, cthreads, ob.smem_col, stream>>>(sa);
            SB_CUDA(cudaGetLastError());
            mark();
            ob.kx.grad<<<rgrid, ob.row_threads, ob.smem_row, stream>>>(sa);
            SB_CUDA(cudaGetLastError());
            mark();
            nk += 5;
        }
        // stage order inside the marks: fwd, kmul, inv, residual, fwd, kmul*, inv (summed over observations)
        for (size_t o = 0; !fused && o < obs.size(); ++o) {
            Obs &ob = *obs[o];
            const DevObs<T> &d = ob.dev;
            const long long per_scene = (long long)d.C * d.Fy * d.Fxc, total = per_scene * S;
            if (d.kind == 0) {
                SB_CUFFT(Cx<T>::fwd(ob.fwd, (typename Cx<T>::real_t *)d.A, (typename Cx<T>::cplx_t *)d.Ahat));
                ++nf;
                mark();
                k_kmul<T><<<grid_for(total, 256, 148 * 32), 256, 0, stream>>>(d.Ahat, d.khat, per_scene, total, d.khat_shared, 0, d_done.p);
                SB_CUDA(cudaGetLastError());
                ++nk;
                mark();
                SB_CUFFT(Cx<T>::inv(ob.inv, (typename Cx<T>::cplx_t *)d.Ahat, (typename Cx<T>::real_t *)d.B));
                ++nf;
                mark();
            } else if (marks) {
                mark(), mark(), mark();
            }
            {
                ResidualArgs<T> ra;
                memset(&ra, 0, sizeof ra);
                ra.ob = d, ra.Ny = desc.Ny, ra.Nx = desc.Nx, ra.done = d_done.p, ra.partials = ob.partials.p;
                ra.rendered_out = ((int)o == rendered_obs) ? rendered_out : nullptr;
                dim3 grid((desc.Nx + 31) / 32, (desc.Ny + 7) / 8, S * d.C), block(32, 8);
                k_residual<T><<<grid, block, 0, stream>>>(ra);
                SB_CUDA(cudaGetLastError());
                ++nk;
            }
            mark();
            if (d.kind == 0) {
                SB_CUFFT(Cx<T>::fwd(ob.fwd, (typename Cx<T>::real_t *)d.A, (typename Cx<T>::cplx_t *)d.Ahat));
                ++nf;
                mark();
                k_kmul<T><<<grid_for(total, 256, 148 * 32), 256, 0, stream>>>(d.Ahat, d.khat, per_scene, total, d.khat_shared, 1, d_done.p);
                SB_CUDA(cudaGetLastError());
                ++nk;
                mark();
                SB_CUFFT(Cx<T>::inv(ob.inv, (typename Cx<T>::cplx_t *)d.Ahat, (typename Cx<T>::real_t *)d.B));
                ++nf;
                mark();
            } else if (marks) {
                mark(), mark(), mark();
            }
        }
        if (n_src) {
            UpdateArgs<T> ua = update_args(mode);
            if (mode == 1 || !use_fast) { // gradients only / forced generic: every source through the generic kernel
                k_update<T><<<n_src, 128, update_smem(), stream>>>(ua);
                SB_CUDA(cudaGetLastError());
                ++nk;
            } else {
                // The three update kernels work on disjoint sources: the warp-per-source kernel goes first on the main stream
                // (one big CTA per SM), the grouped and the generic kernel fill the rest of the machine from side streams
                // (also inside a graph capture).
                const int n_kinds = (n_warp_cta > 0) + (n_fast_cta > 0) + (n_generic > 0);
                const bool fork = n_kinds > 1 && !marks;
                if (fork) SB_CUDA(cudaEventRecord(ev_fork, stream));
                cudaStream_t s_fast = stream, s_gen = stream;
                if (fork && n_warp_cta && n_fast_cta) s_fast = side2;
                if (fork && (n_warp_cta || n_fast_cta) && n_generic) s_gen = side;
                if (n_warp_cta) {
                    WarpArgs<T> wa;
                    wa.groups = d_warp_groups.p, wa.G = warp_G, wa.npix = warp_npix, wa.table_cap = warp_cap, wa.xp = d_xp.p;
                    wa.segs = d_warp_segs.p, wa.counters = d_warp_ctr.p;
                    SB_CUDA(cudaMemsetAsync(d_warp_ctr.p, 0, (size_t)std::max(n_warp_chains, 1) * sizeof(int), stream));
                    wa.use_ring = getenv("SB_NO_XP_RING") == nullptr;
                    wa.bulk_table = getenv("SB_NO_BULK_TABLE") == nullptr;
                    k_update_warp<T, WARP_NPT, WARP_MAXT><<<n_warp_cta, 32 * warp_G, warp_smem, stream>>>(ua, wa);
                    SB_CUDA(cudaGetLastError());
                    ++nk;
                }
                if (n_fast_cta) {
                    if (s_fast != stream) SB_CUDA(cudaStreamWaitEvent(s_fast, ev_fork, 0));
                    if (fast_GT == 32 && 32 * fast_G <= 448)
                        k_update_fast<T, 32, 448><<<n_fast_cta, 32 * fast_G, fast_smem, s_fast>>>(ua);
                    else if (fast_GT == 32)
                        k_update_fast<T, 32><<<n_fast_cta, 32 * fast_G, fast_smem, s_fast>>>(ua);
                    else if (fast_GT == 128)
                        k_update_fast<T, 128><<<n_fast_cta, 128 * fast_G, fast_smem, s_fast>>>(ua);
                    else if (64 * fast_G <= 832)
                        k_update_fast<T, 64, 832><<<n_fast_cta, 64 * fast_G, fast_smem, s_fast>>>(ua);
                    else
                        k_update_fast<T, 64><<<n_fast_cta, 64 * fast_G, fast_smem, s_fast>>>(ua);
                    SB_CUDA(cudaGetLastError());
                    ++nk;
                    if (s_fast != stream) SB_CUDA(cudaEventRecord(ev_join2, s_fast));
                }
                if (n_generic) {
                    if (s_gen != stream) SB_CUDA(cudaStreamWaitEvent(s_gen, ev_fork, 0));
                    UpdateArgs<T> ug = ua;
                    ug.work = d_work.p;
                    k_update<T><<<n_generic, 128, update_smem(), s_gen>>>(ug);
                    SB_CUDA(cudaGetLastError());
                    ++nk;
                    if (s_gen != stream) SB_CUDA(cudaEventRecord(ev_join, s_gen));
                }
                if (s_fast != stream) SB_CUDA(cudaStreamWaitEvent(stream, ev_join2, 0));
                if (s_gen != stream) SB_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
            }
            for (size_t o = 0; fused && o < obs.size(); ++o) { // fitted kernel offsets: gradient, step, new K^ for the next render
                Obs &ob = *obs[o];
                if (!ob.psf_shift) continue;
                PsfShiftArgs<T> pa = psf_args((int)o, mode);
                k_psf_corr<T><<<dim3(ob.Py, ob.dev.C, S), 256, 0, stream>>>(pa);
                k_psf_update<T><<<S, 128, 0, stream>>>(pa);
                SB_CUDA(cudaGetLastError());
                nk += 2;
                if (mode == 0 && !ob.shift_fixed) {
                    SB_TRY(refresh_psf((int)o));
                    nk += 3;
                }
            }
            if (mode == 0 && n_shift) { // new shifts / images -> new shifted morphologies for the next render
                SB_TRY(launch_shift_apply());
                ++nk;
            }
        }
        mark();
        {
            LossArgs la;
            memset(&la, 0, sizeof la);
            la.n_obs = (int)obs.size();
            for (size_t o = 0; o < obs.size(); ++o) la.partials[o] = obs[o]->partials.p, la.n_part[o] = obs[o]->n_part;
            la.loss_const = d_loss_const.p, la.loss = d_loss.p, la.cap = loss_cap, la.done = d_done.p, la.n_iter = d_niter.p;
            la.n_active_next = d_nactive_next.p, la.it_arr = d_it.p, la.status = d_status.p, la.fs = cur_fs;
            la.state = d_state.p, la.limit = use_limit ? d_limit.p : nullptr, la.advance = mode == 0;
            k_loss_stop<<<S, 128, 0, stream>>>(la);
            SB_CUDA(cudaGetLastError());
            ++nk;
            if (mode == 0) {
                k_tick<<<1, 1, 0, stream>>>(d_nactive.p, d_nactive_next.p);
                SB_CUDA(cudaGetLastError());
                ++nk;
            }
        }
        mark();
        kernels_per_iter = nk, ffts_per_iter = nf;
        return SB_OK;
    }

    int ensure_graph(const sb_fit_opts *o) {
        FitScalars f = scalars(o);
        if (have_graph && memcmp(&f, &graph_fs, sizeof f) == 0) return SB_OK;
        cur_fs = f;
        if (graph) {
            cudaGraphExecDestroy(graph);
            graph = nullptr;
        }
        cudaGraph_t g = nullptr;
        SB_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_iteration(0, nullptr, nullptr, -1, nullptr);
        cudaError_t e = cudaStreamEndCapture(stream, &g);
        if (rc != SB_OK) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        if (e != cudaSuccess) return set_err(SB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return set_err(SB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
        graph_fs = f;
        have_graph = true;
        return SB_OK;
    }

    int check_opts(const sb_fit_opts *o) {
        if (!o) return set_err(SB_ERR_ARG, "null fit options");
        if (o->max_iter < 0 || o->prox_max_iter < 0 || o->check_every < 1) return set_err(SB_ERR_ARG, "bad fit options");
        return SB_OK;
    }

    int fit_enqueue(const sb_fit_opts *o, int n) override {
        SB_TRY(check_opts(o));
        SB_CUDA(cudaSetDevice(device));
        SB_TRY(ensure_loss_cap(std::max(n, 1)));
        drop_scene_tables();
        SB_TRY(ensure_graph(o));
        SB_TRY(reset_counters());
        for (int i = 0; i < n; ++i) SB_CUDA(cudaGraphLaunch(graph, stream));
        launches += (int64_t)n * kernels_per_iter;
        return SB_OK;
    }

    int cur_it = 0; // host mirror of the device iteration counter (every graph launch advances it by one)

    int fit(const sb_fit_opts *o, int32_t *n_iter, double *loss, int32_t *status) override {
        SB_TRY(check_opts(o));
        SB_CUDA(cudaSetDevice(device));
        SB_TRY(ensure_loss_cap(std::max(o->max_iter, 1)));
        if (!o->resume) drop_scene_tables();
        SB_TRY(ensure_graph(o));
        if (!o->resume) {
            SB_TRY(reset_counters());
            cur_it = 0;
        }
        const int end = o->run_until > 0 ? std::min(o->run_until, o->max_iter) : o->max_iter;
        int done_iters = 0;
        while (cur_it < end) {
            SB_CUDA(cudaGraphLaunch(graph, stream));
            ++done_iters, ++cur_it;
            if (!o->fixed_iterations && (cur_it % o->check_every == 0)) {
                SB_CUDA(cudaMemcpyAsync(h_nactive, d_nactive.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
                SB_CUDA(cudaStreamSynchronize(stream));
                if (*h_nactive == 0) break;
            }
        }
        launches += (int64_t)done_iters * kernels_per_iter;
        if (n_iter) SB_CUDA(cudaMemcpyAsync(n_iter, d_niter.p, S * sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (status) SB_CUDA(cudaMemcpyAsync(status, d_status.p, S * sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (loss && o->max_iter > 0)
            SB_CUDA(cudaMemcpy2DAsync(loss, (size_t)o->max_iter * sizeof(double), d_loss.p, (size_t)loss_cap * sizeof(double),
                                      (size_t)o->max_iter * sizeof(double), S, cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }

    // ---- per-scene control of a running fit (batches whose scenes restart independently: dynamic boxes) ------------------
    // Every scene carries its own adaprox iteration counter, loss-history length, iteration budget and run flag; `run`
    // launches iterations until no scene is running any more (converged, paused for inspection, budget spent, failed).
    int scene_control(const int32_t *it_local, const int32_t *loss_len, const int32_t *limit, const int32_t *active, const int32_t *prox_iter) override {
        SB_CUDA(cudaSetDevice(device));
        if (it_local) SB_CUDA(cudaMemcpyAsync(d_it.p, it_local, S * sizeof(int), cudaMemcpyHostToDevice, stream));
        if (loss_len) SB_CUDA(cudaMemcpyAsync(d_niter.p, loss_len, S * sizeof(int), cudaMemcpyHostToDevice, stream));
        if (limit) {
            SB_CUDA(cudaMemcpyAsync(d_limit.p, limit, S * sizeof(int), cudaMemcpyHostToDevice, stream));
            if (!use_limit) have_graph = false;
            use_limit = true;
        }
        if (prox_iter) {
            SB_CUDA(cudaMemcpyAsync(d_prox_iter.p, prox_iter, S * sizeof(int), cudaMemcpyHostToDevice, stream));
            if (!use_prox_iter) have_graph = false;
            use_prox_iter = true;
        }
        if (active) {
            std::vector<int> done(S), state(S);
            int n = 0;
            for (int s = 0; s < S; ++s) done[s] = active[s] ? 0 : 1, n += active[s] ? 1 : 0;
            SB_CUDA(cudaMemcpyAsync(d_done.p, done.data(), S * sizeof(int), cudaMemcpyHostToDevice, stream));
            SB_CUDA(cudaMemcpyAsync(d_nactive.p, &n, sizeof(int), cudaMemcpyHostToDevice, stream));
            SB_TRY(d_nactive_next.zero(stream));
            SB_TRY(d_status.zero(stream));
            SB_CUDA(cudaStreamSynchronize(stream)); // host vectors die here
        }
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }
    int scene_status(int32_t *it_local, int32_t *loss_len, int32_t *state) override {
        SB_CUDA(cudaSetDevice(device));
        if (it_local) SB_CUDA(cudaMemcpyAsync(it_local, d_it.p, S * sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (loss_len) SB_CUDA(cudaMemcpyAsync(loss_len, d_niter.p, S * sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (state) SB_CUDA(cudaMemcpyAsync(state, d_state.p, S * sizeof(int), cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }
    int run(const sb_fit_opts *o, int max_launches, int32_t *launched) override {
        SB_TRY(check_opts(o));
        SB_CUDA(cudaSetDevice(device));
        SB_TRY(ensure_loss_cap(std::max(o->max_iter, 1)));
        SB_TRY(ensure_graph(o));
        int n = 0;
        while (n < max_launches) {
            SB_CUDA(cudaGraphLaunch(graph, stream));
            ++n;
            if (n % o->check_every == 0 || n == max_launches) {
                SB_CUDA(cudaMemcpyAsync(h_nactive, d_nactive.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
                SB_CUDA(cudaStreamSynchronize(stream));
                if (*h_nactive == 0) break;
            }
        }
        launches += (int64_t)n * kernels_per_iter;
        if (launched) *launched = n;
        return SB_OK;
    }
    int download_loss(double *loss, int n_cols) override {
        SB_CUDA(cudaSetDevice(device));
        if (!loss || n_cols <= 0 || n_cols > loss_cap) return set_err(SB_ERR_ARG, "download_loss: %d columns requested, capacity %d", n_cols, loss_cap);
        SB_CUDA(cudaMemcpy2DAsync(loss, (size_t)n_cols * sizeof(double), d_loss.p, (size_t)loss_cap * sizeof(double),
                                  (size_t)n_cols * sizeof(double), S, cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }

    int upload_loss(const double *loss, int n_cols) override {
        SB_CUDA(cudaSetDevice(device));
        if (!loss || n_cols <= 0) return set_err(SB_ERR_ARG, "upload_loss: bad argument");
        SB_TRY(ensure_loss_cap(n_cols));
        SB_CUDA(cudaMemcpy2DAsync(d_loss.p, (size_t)loss_cap * sizeof(double), loss, (size_t)n_cols * sizeof(double),
                                  (size_t)n_cols * sizeof(double), S, cudaMemcpyHostToDevice, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }

    int profile(const sb_fit_opts *o, int n, float *stage_ms) override {
        SB_TRY(check_opts(o));
        SB_CUDA(cudaSetDevice(device));
        SB_TRY(ensure_loss_cap(std::max(n, 1)));
        cur_fs = scalars(o);
        have_graph = false;
        drop_scene_tables();
        SB_TRY(reset_counters());
        for (int i = 0; i < SB_N_STAGES; ++i) stage_ms[i] = 0.f;
        for (int i = 0; i < n; ++i) {
            std::vector<cudaEvent_t> marks;
            SB_TRY(enqueue_iteration(0, nullptr, nullptr, -1, &marks));
            SB_CUDA(cudaStreamSynchronize(stream));
            // marks: [start, after render, (7 per observation), after update, after advance]
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[0], marks[1]);
            stage_ms[0] += ms;
            size_t idx = 1;
            for (size_t ob = 0; ob < obs.size(); ++ob)
                for (int st = 1; st <= 7; ++st) {
                    cudaEventElapsedTime(&ms, marks[idx], marks[idx + 1]);
                    stage_ms[st] += ms;
                    ++idx;
                }
            cudaEventElapsedTime(&ms, marks[idx], marks[idx + 1]);
            stage_ms[8] += ms;
            cudaEventElapsedTime(&ms, marks[idx + 1], marks[idx + 2]);
            stage_ms[9] += ms;
            for (auto e : marks) cudaEventDestroy(e);
        }
        launches += (int64_t)n * kernels_per_iter;
        return SB_OK;
    }

    int evaluate(int o, double *model, double *rendered, double *loss, double *g_sed, double *g_morph, double *g_center) override {
        SB_CUDA(cudaSetDevice(device));
        if (o < 0 || o >= (int)obs.size()) return set_err(SB_ERR_ARG, "observation index %d out of range", o);
        SB_TRY(ensure_loss_cap(1));
        drop_scene_tables();
        SB_TRY(reset_counters());
        const size_t nmodel = (size_t)S * C * desc.Ny * desc.Nx, nrend = obs[o]->data.n;
        if (model) {
            if (d_model.n < nmodel) SB_TRY(d_model.alloc(nmodel));
            SB_CUDA(cudaMemsetAsync(d_model.p, 0, nmodel * sizeof(T), stream));
        }
        if (rendered) {
            if (d_rendered.n < nrend) SB_TRY(d_rendered.alloc(nrend));
            SB_CUDA(cudaMemsetAsync(d_rendered.p, 0, nrend * sizeof(T), stream));
        }
        if (g_morph && d_gmorph.n < (size_t)std::max<long long>(n_morph, 1)) SB_TRY(d_gmorph.alloc(std::max<long long>(n_morph, 1)));
        T *keep_rendered = rendered ? d_rendered.p : nullptr;
        double *keep_gm = d_gmorph.p;
        if (!g_morph) d_gmorph.p = nullptr; // kernel skips the write
        int rc = enqueue_iteration(1, model ? d_model.p : nullptr, keep_rendered, o, nullptr);
        d_gmorph.p = keep_gm;
        SB_TRY(rc);
        launches += kernels_per_iter;
        SB_TRY(download_T(model, d_model.p, model ? nmodel : 0));
        SB_TRY(download_T(rendered, d_rendered.p, rendered ? nrend : 0));
        if (loss) SB_CUDA(cudaMemcpy2DAsync(loss, sizeof(double), d_loss.p, (size_t)loss_cap * sizeof(double), sizeof(double), S, cudaMemcpyDeviceToHost, stream));
        if (g_sed && n_src) SB_CUDA(cudaMemcpyAsync(g_sed, d_gsed.p, (size_t)n_src * C * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (g_center && n_point) SB_CUDA(cudaMemcpyAsync(g_center, d_gcenter.p, (size_t)n_point * 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        if (g_morph && n_morph) SB_CUDA(cudaMemcpyAsync(g_morph, d_gmorph.p, (size_t)n_morph * sizeof(double), cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
        return SB_OK;
    }
};

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char *sb_last_error(void) { return g_err.c_str(); }
const char *sb_version(void) { return "scarlet_b200 0.2 (sm_100a)"; }
#ifndef SB_SRC_HASH
#define SB_SRC_HASH "unknown"
#endif
const char *sb_source_hash(void) { return SB_SRC_HASH; }
int sb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
const char *sb_stage_name(int s) { return (s >= 0 && s < SB_N_STAGES) ? kStageNames[s] : ""; }

static int need_device(int device) {
    int n = sb_device_count();
    if (n <= 0) return set_err(SB_ERR_CUDA, "no CUDA device available: scarlet_b200 has no CPU fallback");
    if (device < 0 || device >= n) return set_err(SB_ERR_ARG, "device %d out of range (have %d)", device, n);
    return SB_OK;
}

int sb_plan_create(const sb_batch_desc *desc, int device, sb_plan **out) {
    if (!desc || !out) return set_err(SB_ERR_ARG, "null argument");
    *out = nullptr;
    SB_TRY(need_device(device));
    sb_plan *p = nullptr;
    int rc;
    if (desc->precision == 32) {
        auto *q = new PlanT<float>();
        rc = q->init(desc, device);
        p = q;
    } else if (desc->precision == 64) {
        auto *q = new PlanT<double>();
        rc = q->init(desc, device);
        p = q;
    } else
        return set_err(SB_ERR_ARG, "precision must be 32 or 64");
    if (rc != SB_OK) {
        delete p;
        return rc;
    }
    *out = p;
    return SB_OK;
}
void sb_plan_destroy(sb_plan *plan) {
    if (plan) {
        cudaSetDevice(plan->device);
        cudaStreamSynchronize(plan->stream);
        delete plan;
    }
}
int64_t sb_plan_device_bytes(const sb_plan *plan) { return plan ? plan->dev_bytes : 0; }
int64_t sb_plan_kernel_launches(const sb_plan *plan) { return plan ? plan->launches : 0; }
void *sb_plan_stream(sb_plan *plan) { return plan ? (void *)plan->stream : nullptr; }

void *sb_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void sb_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

#define PLAN_CALL(expr)                                   \
    if (!plan) return set_err(SB_ERR_ARG, "null plan");   \
    return plan->expr;

int sb_plan_upload_observation(sb_plan *plan, int obs, const float *data, const float *weights, const double *khat, const double *loss_const) {
    PLAN_CALL(upload_observation(obs, data, weights, 4, khat, loss_const))
}
int sb_plan_upload_observation_f64(sb_plan *plan, int obs, const double *data, const double *weights, const double *khat, const double *loss_const) {
    PLAN_CALL(upload_observation(obs, data, weights, 8, khat, loss_const))
}
int sb_plan_upload_kernels(sb_plan *plan, int obs, const double *kernels, int Py, int Px, int y0, int x0) {
    PLAN_CALL(upload_kernels(obs, kernels, Py, Px, y0, x0))
}
int sb_plan_zero_state(sb_plan *plan) { PLAN_CALL(zero_state()) }
int sb_plan_upload_resampling(sb_plan *plan, int obs, const double *ey, const double *ex, double h2) {
    PLAN_CALL(upload_resampling(obs, ey, ex, h2))
}
int sb_plan_upload_resampling_rot(sb_plan *plan, int obs, const double *a, const double *b, double h2) {
    PLAN_CALL(upload_resampling_rot(obs, a, b, h2))
}
int sb_host_gather_f64(double *dst, const void *const *src, const int64_t *count, const int32_t *is_f32, int64_t n) {
    if (!dst || !src || !count || !is_f32 || n < 0) return set_err(SB_ERR_ARG, "null argument");
    for (int64_t i = 0; i < n; ++i) {
        const int64_t c = count[i];
        if (is_f32[i]) {
            const float *p = static_cast<const float *>(src[i]);
            for (int64_t j = 0; j < c; ++j) dst[j] = (double)p[j];
        } else
            memcpy(dst, src[i], (size_t)c * sizeof(double));
        dst += c;
    }
    return SB_OK;
}
int sb_host_scatter_f64(const double *src, void *const *dst, const int64_t *count, const int32_t *is_f32, int64_t n) {
    if (!dst || !src || !count || !is_f32 || n < 0) return set_err(SB_ERR_ARG, "null argument");
    for (int64_t i = 0; i < n; ++i) {
        const int64_t c = count[i];
        if (is_f32[i]) {
            float *p = static_cast<float *>(dst[i]);
            for (int64_t j = 0; j < c; ++j) p[j] = (float)src[j];
        } else
            memcpy(dst[i], src, (size_t)c * sizeof(double));
        src += c;
    }
    return SB_OK;
}
int sb_plan_upload_params(sb_plan *plan, int which, const double *sed, const double *morph, const double *center) {
    PLAN_CALL(upload_params(which, sed, morph, center))
}
int sb_plan_download_params(sb_plan *plan, int which, double *sed, double *morph, double *center) {
    PLAN_CALL(download_params(which, sed, morph, center))
}
int sb_plan_evaluate(sb_plan *plan, int obs, double *model, double *rendered, double *loss, double *g_sed, double *g_morph, double *g_center) {
    PLAN_CALL(evaluate(obs, model, rendered, loss, g_sed, g_morph, g_center))
}
int sb_plan_fit(sb_plan *plan, const sb_fit_opts *opts, int32_t *n_iter_out, double *loss_out, int32_t *status_out) {
    PLAN_CALL(fit(opts, n_iter_out, loss_out, status_out))
}
int sb_plan_fit_enqueue(sb_plan *plan, const sb_fit_opts *opts, int n_iterations) { PLAN_CALL(fit_enqueue(opts, n_iterations)) }
int sb_plan_profile_iterations(sb_plan *plan, const sb_fit_opts *opts, int n_iterations, float *stage_ms) {
    if (!stage_ms) return set_err(SB_ERR_ARG, "null stage_ms");
    PLAN_CALL(profile(opts, n_iterations, stage_ms))
}
int sb_plan_device_params(sb_plan *plan, void **sed, int64_t *n_sed, void **morph, int64_t *n_morph, int *elem_bytes) {
    PLAN_CALL(device_params(sed, n_sed, morph, n_morph, elem_bytes))
}
int sb_plan_spectral_mode(const sb_plan *plan) { return plan ? plan->spectral_mode() : -1; }
int sb_plan_prox_histogram(sb_plan *plan, int enable, int64_t *out16) { PLAN_CALL(prox_histogram(enable, out16)) }
int sb_plan_scene_control(sb_plan *plan, const int32_t *it_local, const int32_t *loss_len, const int32_t *limit, const int32_t *active,
                          const int32_t *prox_iter) {
    PLAN_CALL(scene_control(it_local, loss_len, limit, active, prox_iter))
}
int sb_plan_scene_status(sb_plan *plan, int32_t *it_local, int32_t *loss_len, int32_t *state) { PLAN_CALL(scene_status(it_local, loss_len, state)) }
int sb_plan_run(sb_plan *plan, const sb_fit_opts *opts, int max_launches, int32_t *launched) { PLAN_CALL(run(opts, max_launches, launched)) }
int sb_plan_download_loss(sb_plan *plan, double *loss, int n_cols) { PLAN_CALL(download_loss(loss, n_cols)) }
int sb_plan_upload_loss(sb_plan *plan, const double *loss, int n_cols) { PLAN_CALL(upload_loss(loss, n_cols)) }
int sb_plan_inspect(sb_plan *plan, int32_t *action) { PLAN_CALL(inspect(action)) }
int sb_plan_set_sources(sb_plan *plan, const sb_batch_desc *desc) {
    if (!desc) return set_err(SB_ERR_ARG, "null descriptor");
    PLAN_CALL(set_sources(desc))
}
int sb_fft_supported_length(int need) { return spec_supported_length(need); }
int sb_plan_sync(sb_plan *plan) {
    if (!plan) return set_err(SB_ERR_ARG, "null plan");
    SB_CUDA(cudaSetDevice(plan->device));
    SB_CUDA(cudaStreamSynchronize(plan->stream));
    return SB_OK;
}
int sb_plan_timer_start(sb_plan *plan) {
    if (!plan) return set_err(SB_ERR_ARG, "null plan");
    SB_CUDA(cudaEventRecord(plan->ev0, plan->stream));
    return SB_OK;
}
int sb_plan_timer_stop(sb_plan *plan, float *ms) {
    if (!plan || !ms) return set_err(SB_ERR_ARG, "null argument");
    SB_CUDA(cudaEventRecord(plan->ev1, plan->stream));
    SB_CUDA(cudaEventSynchronize(plan->ev1));
    SB_CUDA(cudaEventElapsedTime(ms, plan->ev0, plan->ev1));
    return SB_OK;
}

} // extern "C"

// ---- single-operator entry points -------------------------------------------------------------------
template <typename T>
static int run_chain(T *img, int By, int Bx, int n_img, const sb_chain_desc *chain, const sb_mono_desc *mono, int n_mono, int device) {
    if (!img || !chain || By <= 0 || Bx <= 0 || n_img <= 0) return set_err(SB_ERR_ARG, "bad argument");
    SB_TRY(need_device(device));
    SB_TRY(check_chain(*chain, n_mono));
    SB_CUDA(cudaSetDevice(device));
    const int n = By * Bx;
    if (n >= 0xffff) return set_err(SB_ERR_ARG, "image too large (%d px)", n);
    std::vector<std::unique_ptr<MonoDevice<T>>> monos;
    std::vector<DevMono> hm(std::max(n_mono, 1));
    for (int i = 0; i < n_mono; ++i) {
        if (mono[i].n_pix != n) return set_err(SB_ERR_ARG, "monotonic table %d has %d pixels, image has %d", i, mono[i].n_pix, n);
        HostMono h;
        SB_TRY(build_mono<T>(mono[i], h));
        monos.emplace_back(new MonoDevice<T>());
        SB_TRY(monos.back()->upload(h, 0));
        hm[i] = monos.back()->dev;
    }
    DevBuf<DevMono> dmo;
    DevBuf<DevChain> dch;
    DevBuf<T> dimg;
    SB_TRY(dmo.alloc(hm.size()));
    SB_TRY(dch.alloc(1));
    SB_TRY(dimg.alloc((size_t)n * n_img));
    DevChain hc;
    hc.n_ops = chain->n_ops, hc.repeat = chain->repeat;
    memcpy(hc.ops, chain->ops, sizeof hc.ops);
    SB_CUDA(cudaMemcpy(dmo.p, hm.data(), hm.size() * sizeof(DevMono), cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(dch.p, &hc, sizeof hc, cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(dimg.p, img, (size_t)n * n_img * sizeof(T), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(((n + 1) & ~1) + 2) * sizeof(T) + 48 * sizeof(double);
    SB_TRY(raise_smem_limit(device, (const void *)k_chain_only<T>, smem));
    k_chain_only<T><<<n_img, 128, smem>>>(dimg.p, By, Bx, dch.p, dmo.p);
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaMemcpy(img, dimg.p, (size_t)n * n_img * sizeof(T), cudaMemcpyDeviceToHost));
    return SB_OK;
}

template <typename T>
static int run_monotonic(T *flat_img, const T *weights, const int32_t *offsets, int n_off, const int32_t *dist_idx, int n_idx,
                         int n_pix, T min_gradient, int n_img, int device) {
    if (!flat_img || !weights || !offsets || n_pix <= 0) return set_err(SB_ERR_ARG, "bad argument");
    std::vector<double> w64((size_t)n_off * n_pix);
    for (size_t i = 0; i < w64.size(); ++i) w64[i] = (double)weights[i];
    sb_mono_desc md;
    memset(&md, 0, sizeof md);
    md.n_pix = n_pix, md.n_off = n_off, md.n_idx = n_idx, md.weights = w64.data(), md.offsets = offsets, md.dist_idx = dist_idx;
    sb_chain_desc ch;
    memset(&ch, 0, sizeof ch);
    ch.n_ops = 1, ch.repeat = 1;
    ch.ops[0].code = SB_OP_MONOTONIC, ch.ops[0].iarg = 0, ch.ops[0].farg = (double)min_gradient;
    // the image is 1-D for this operator: treat it as a single row
    return run_chain<T>(flat_img, 1, n_pix, n_img, &ch, &md, 1, device);
}

extern "C" {

int sb_monotonic_f32(float *flat_img, const float *weights, const int32_t *offsets, int n_off, const int32_t *dist_idx, int n_idx,
                     int n_pix, float min_gradient, int n_img, int device) {
    return run_monotonic<float>(flat_img, weights, offsets, n_off, dist_idx, n_idx, n_pix, min_gradient, n_img, device);
}
int sb_monotonic_f64(double *flat_img, const double *weights, const int32_t *offsets, int n_off, const int32_t *dist_idx, int n_idx,
                     int n_pix, double min_gradient, int n_img, int device) {
    return run_monotonic<double>(flat_img, weights, offsets, n_off, dist_idx, n_idx, n_pix, min_gradient, n_img, device);
}
int sb_prox_chain_f32(float *img, int By, int Bx, int n_img, const sb_chain_desc *chain, const sb_mono_desc *mono, int n_mono, int device) {
    return run_chain<float>(img, By, Bx, n_img, chain, mono, n_mono, device);
}
int sb_prox_chain_f64(double *img, int By, int Bx, int n_img, const sb_chain_desc *chain, const sb_mono_desc *mono, int n_mono, int device) {
    return run_chain<double>(img, By, Bx, n_img, chain, mono, n_mono, device);
}

} // extern "C"

template <typename T>
static int run_convolve(const T *image, int C, int Ny, int Nx, const double *khat, int Fy, int Fx, int adjoint, T *out, int device) {
    typedef typename Cx<T>::type cplx;
    if (!image || !khat || !out || C <= 0 || Ny <= 0 || Nx <= 0 || Fy < Ny || Fx < Nx) return set_err(SB_ERR_ARG, "bad argument");
    SB_TRY(need_device(device));
    SB_CUDA(cudaSetDevice(device));
    const int Fxc = Fx / 2 + 1;
    const size_t nimg = (size_t)C * Ny * Nx, ngrid = (size_t)C * Fy * Fx, nc = (size_t)C * Fy * Fxc;
    DevBuf<T> dimg, A, B;
    DevBuf<cplx> Ahat, K;
    DevBuf<double2> ks;
    SB_TRY(dimg.alloc(nimg));
    SB_TRY(A.alloc(ngrid));
    SB_TRY(B.alloc(ngrid));
    SB_TRY(Ahat.alloc(nc));
    SB_TRY(K.alloc(nc));
    SB_TRY(ks.alloc(nc));
    SB_TRY(A.zero(0));
    SB_CUDA(cudaMemcpy(dimg.p, image, nimg * sizeof(T), cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(ks.p, khat, nc * sizeof(double2), cudaMemcpyHostToDevice));
    k_cast_scale_cplx<T><<<grid_for(nc), 256>>>(ks.p, K.p, (long long)nc, 1.0 / ((double)Fy * Fx));
    k_embed<T><<<grid_for(nimg), 256>>>(dimg.p, A.p, C, Fy, Fx, Ny, Nx);
    SB_CUDA(cudaGetLastError());
    cufftHandle fwd = 0, inv = 0;
    int n[2] = {Fy, Fx};
    SB_CUFFT(cufftPlanMany(&fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, Cx<T>::r2c, C));
    cufftResult r2 = cufftPlanMany(&inv, 2, n, nullptr, 1, 0, nullptr, 1, 0, Cx<T>::c2r, C);
    if (r2 != CUFFT_SUCCESS) {
        cufftDestroy(fwd);
        return set_err(SB_ERR_CUFFT, "cufftPlanMany(c2r) -> %d", (int)r2);
    }
    cufftResult r = Cx<T>::fwd(fwd, (typename Cx<T>::real_t *)A.p, (typename Cx<T>::cplx_t *)Ahat.p);
    if (r == CUFFT_SUCCESS) {
        k_kmul<T><<<grid_for(nc), 256>>>(Ahat.p, K.p, (long long)nc, (long long)nc, 1, adjoint, nullptr);
        r = Cx<T>::inv(inv, (typename Cx<T>::cplx_t *)Ahat.p, (typename Cx<T>::real_t *)B.p);
    }
    cufftDestroy(fwd);
    cufftDestroy(inv);
    if (r != CUFFT_SUCCESS) return set_err(SB_ERR_CUFFT, "cufft exec -> %d", (int)r);
    k_crop<T><<<grid_for(nimg), 256>>>(B.p, dimg.p, C, Fy, Fx, Ny, Nx);
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaMemcpy(out, dimg.p, nimg * sizeof(T), cudaMemcpyDeviceToHost));
    return SB_OK;
}
extern "C" {
int sb_fft_convolve_f32(const float *image, int C, int Ny, int Nx, const double *khat, int Fy, int Fx, int adjoint, float *out, int device) {
    return run_convolve<float>(image, C, Ny, Nx, khat, Fy, Fx, adjoint, out, device);
}
int sb_fft_convolve_f64(const double *image, int C, int Ny, int Nx, const double *khat, int Fy, int Fx, int adjoint, double *out, int device) {
    return run_convolve<double>(image, C, Ny, Nx, khat, Fy, Fx, adjoint, out, device);
}

} // extern "C"

// ---- apply_filter ------------------------------------------------------------------------------------------------
template <typename T>
static int run_apply_filter(const T *image, int H, int W, const T *values, const int32_t *ys, const int32_t *ye, const int32_t *xs,
                            const int32_t *xe, int n_taps, T *result, int device) {
    if (!image || !values || !ys || !ye || !xs || !xe || !result || H <= 0 || W <= 0 || n_taps < 0) return set_err(SB_ERR_ARG, "bad argument");
    SB_TRY(need_device(device));
    SB_CUDA(cudaSetDevice(device));
    DevBuf<T> dimg, dval, dout;
    DevBuf<int> db[4];
    const int32_t *hb[4] = {ys, ye, xs, xe};
    SB_TRY(dimg.alloc((size_t)H * W));
    SB_TRY(dout.alloc((size_t)H * W));
    SB_TRY(dval.alloc(std::max(n_taps, 1)));
    SB_CUDA(cudaMemcpy(dimg.p, image, (size_t)H * W * sizeof(T), cudaMemcpyHostToDevice));
    if (n_taps) SB_CUDA(cudaMemcpy(dval.p, values, (size_t)n_taps * sizeof(T), cudaMemcpyHostToDevice));
    for (int i = 0; i < 4; ++i) {
        SB_TRY(db[i].alloc(std::max(n_taps, 1)));
        if (n_taps) SB_CUDA(cudaMemcpy(db[i].p, hb[i], (size_t)n_taps * sizeof(int), cudaMemcpyHostToDevice));
    }
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8);
    k_apply_filter<T><<<grid, block>>>(dimg.p, H, W, dval.p, db[0].p, db[1].p, db[2].p, db[3].p, n_taps, dout.p);
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaMemcpy(result, dout.p, (size_t)H * W * sizeof(T), cudaMemcpyDeviceToHost));
    return SB_OK;
}
extern "C" {
int sb_apply_filter_f32(const float *image, int H, int W, const float *values, const int32_t *y_start, const int32_t *y_end,
                        const int32_t *x_start, const int32_t *x_end, int n_taps, float *result, int device) {
    return run_apply_filter<float>(image, H, W, values, y_start, y_end, x_start, x_end, n_taps, result, device);
}
int sb_apply_filter_f64(const double *image, int H, int W, const double *values, const int32_t *y_start, const int32_t *y_end,
                        const int32_t *x_start, const int32_t *x_end, int n_taps, double *result, int device) {
    return run_apply_filter<double>(image, H, W, values, y_start, y_end, x_start, x_end, n_taps, result, device);
}
} // extern "C"
