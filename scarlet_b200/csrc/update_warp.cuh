// K4-warp  per-source backward + AMSGrad adaprox update + constraint projections, ONE WARP PER SOURCE (sm_100a).
//
// The throughput path for image morphologies with the ExtendedSource constraint chain
//     Monotonicity -> [Symmetry] -> Positivity -> CenterOn -> Normalization("max")     (morphology.py:644-669)
// on odd boxes of up to 32*NPT pixels.  Reference arithmetic: gradients = what autograd.grad yields at blend.py:118,
// update = proxmin.adaprox(scheme="amsgrad", prox_max_iter) as called at blend.py:165-180, projections
// constraint.py:83-114, 183-287 and operators_pybind11.cc:14-36 -- the same formulas as update_extended (kernels.cuh).
//
// Why one warp: the proximal loop runs up to 10 sub-iterations of a ~60-level radial wavefront per source and iteration.
// With a source per warp the levels are separated by __syncwarp only (no named barriers, no idle partner warp), every
// reduction is a shuffle, and the running iterate z lives in REGISTERS (NPT pixels per lane), so that one proximal
// sub-iteration touches global memory only to stream the gradient-step result x and the metric psi (8 bytes per pixel,
// interleaved, L2-resident).  All warps of a CTA share the wavefront table of their common constraint chain in shared
// memory; levels wider than a warp are split into trips of <= 32 tasks when the table is staged, and the table entries of
// trip t+1 are fetched while trip t computes.
//
// Float path: the element-wise arithmetic of the proximal loop runs in float (the iterate is stored in float anyway, and
// the convergence sums see a relative noise of ~1e-4 from that storage alone, far above float accumulation error); the
// gradient contraction over bands stays in double.  Double path (parity twin): everything in double, true division.
#pragma once
#include "kernels.cuh"

namespace sb {

template <typename T> struct XP { T x, psi; }; // gradient-step result and AMSGrad metric of one pixel, interleaved

// ---- bulk asynchronous copies (the copy engine behind TMA: cp.async.bulk, SASS UBLKCP) + mbarrier completion ----------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
// generic-proxy writes (st.global / st.shared) -> visible to the asynchronous proxy that executes the bulk copies
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// x / psi of one source travel from global memory (L2) to the proximal loop through a small shared-memory ring that a
// bulk asynchronous copy fills XP_AHEAD chunks ahead of the consumer: no registers, no exposed L2 latency.
template <typename T> struct XpRing {
    static constexpr int TRIPS = 4;                             // trips (of 32 pixels) per chunk = the load batch of pass B
    static constexpr int STAGES = 2;                            // power of two; 2 KB per warp keeps 14 warps per CTA at 41 x 41
    static constexpr int CHUNK = TRIPS * 32 * (int)sizeof(XP<T>); // bytes per chunk
    static constexpr int BYTES = STAGES * CHUNK;
};

template <typename T> struct WarpArgs {
    const int *groups;   // source lists of the chains, back to back
    const int4 *segs;    // [n_cta] {offset of the CTA's chain list in groups, its length, index of the chain's counter, -}
    int *counters;       // [n_chains] next unclaimed entry of each list (zeroed before every launch)
    int G;               // warps per CTA
    int npix;            // shared-memory image length per warp (largest box + spare cell, padded)
    int table_cap;       // task capacity of the shared-memory table
    XP<T> *xp;           // one padded slot of 32 NPT entries per warp
    int bulk_table;      // 1: operator table by bulk asynchronous copy; 0: copied by all threads (diagnostic switch)
    int use_ring;        // 1: x / psi reach pass B through the bulk-copy ring; 0: plain global loads (diagnostic switch)
};

__device__ __forceinline__ float warp_sum_all_t(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_all_t(double v) { return warp_sum_all(v); }
template <typename T> __device__ __forceinline__ T warp_max_all_t(T a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T b = __shfl_xor_sync(0xffffffffu, a, o);
        a = b > a ? b : a;
    }
    return a;
}

// AMSGrad in the working precision of the state arrays (the double twin calls amsgrad() of kernels.cuh)
__device__ __forceinline__ float amsgrad_t(float g, float &m, float &v, float &vhat, int it, const FitScalars &fs) {
    // the complements are formed in double: 1.f - (float)0.999 is off by 1.3e-5 relative, a systematic scale error of v
    const float b1 = (float)fs.b1, b2 = (float)fs.b2, omb1 = (float)(1.0 - fs.b1), omb2 = (float)(1.0 - fs.b2), eps = (float)fs.eps;
    m = omb1 * g + b1 * m;
    v = omb2 * (g * g) + b2 * v;
    vhat = (it == 0 && fs.overwrite_vhat_at_it0) ? v : fmaxf(vhat, v);
    return sqrtf(eps > 0.f ? fmaxf(vhat, eps) : vhat);
}
__device__ __forceinline__ double amsgrad_t(double g, double &m, double &v, double &vhat, int it, const FitScalars &fs) {
    return amsgrad(g, m, v, vhat, it, fs);
}

// Shared-memory image of a DevMono with nb == 4, laid out by TRIPS: every dependency level is cut into trips of exactly 32
// slots (entry 32 t + lane); slots beyond the level's tasks hold a dummy task (own pixel and all neighbours = the spare
// zero cell behind the image, weights 0), which reads 0, computes cap = 0 and never stores (0 < 0 is false).  The sweep
// therefore runs without a single predicate or divergent branch, and its table addresses advance by constants.
template <typename T> struct WarpTable {
    unsigned nbr, w, pix; // 32-bit shared-window addresses of uint2[32 n_trips], W4<T>[32 n_trips], u16[32 n_trips], already + lane
    int n_trips;          // even (a dummy trip is appended when needed)
};

template <typename T> struct SweepRec {
    uint2 nb;
    W4<T> w;
    unsigned pp;
};
template <typename T> __device__ __forceinline__ void sweep_fetch(SweepRec<T> &r, const WarpTable<T> &t, unsigned trip) {
    r.nb = lds_u32x2(t.nbr + 256u * trip);
    r.w = lds_w4(t.w + 32u * (unsigned)sizeof(W4<T>) * trip, T(0));
    r.pp = lds_u16(t.pix + 64u * trip);
}
template <typename T> __device__ __forceinline__ void sweep_apply(unsigned zb, const SweepRec<T> &r, T keep) {
    const unsigned ap = zb + r.pp;
    const T v0 = lds_real(zb + (r.nb.x & 0xffffu), T(0)), v1 = lds_real(zb + (r.nb.x >> 16), T(0));
    const T v2 = lds_real(zb + (r.nb.y & 0xffffu), T(0)), v3 = lds_real(zb + (r.nb.y >> 16), T(0));
    const T own = lds_real(ap, T(0));
    T ref = mul_rn(v0, r.w.a);
    ref = add_rn(ref, mul_rn(v1, r.w.b));
    ref = add_rn(ref, mul_rn(v2, r.w.c));
    ref = add_rn(ref, mul_rn(v3, r.w.d));
    const T cap = mul_rn(ref, keep);
    if (cap < own) sts_real(ap, cap);
}

// One sweep of the radial-monotonicity projection over the image at shared address zb (operators_pybind11.cc:14-36).
// Tasks of one trip are independent (same dependency level); __syncwarp orders the trips.  The table entries of the next
// trip do not depend on the image: they are fetched before the current trip's arithmetic (two register sets, no moves).
template <typename T> __device__ __forceinline__ void warp_sweep(unsigned zb, const WarpTable<T> &t, T keep) {
    SweepRec<T> ra, rb;
    sweep_fetch<T>(ra, t, 0u);
#pragma unroll 1
    for (int trip = 0; trip < t.n_trips; trip += 2) {
        sweep_fetch<T>(rb, t, (unsigned)trip + 1u);
        sweep_apply<T>(zb, ra, keep);
        __syncwarp();
        sweep_fetch<T>(ra, t, (unsigned)trip + 2u); // the table carries one more (dummy) trip behind the last
        sweep_apply<T>(zb, rb, keep);
        __syncwarp();
    }
}

// NPT: pixels per lane (register-resident iterate); MAXT: upper bound of the CTA size (sets the register budget)
template <typename T, int NPT, int MAXT> __global__ void __launch_bounds__(MAXT, 1) k_update_warp(const UpdateArgs<T> a, const WarpArgs<T> wa) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int G = wa.G, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The CTA serves one constraint chain (one operator table); its warps claim sources of that chain one at a time from a
    // counter shared by all CTAs of the chain, so a warp whose source converged after two proximal sub-iterations moves on
    // while its neighbour runs all ten: the kernel ends when the work is done, not when the unluckiest CTA is.
    const int4 seg = wa.segs[blockIdx.x];
    const int *list = wa.groups + seg.x;
    const int n_list = seg.y;
    int *counter = wa.counters + seg.z;
    // ---- shared memory: table by trips (W4 | uint2 | u16 pix, table_cap entries each), barriers, then per warp: spectrum
    // gradient, image, x/psi ring
    typedef XpRing<T> Ring;
    const int cap = wa.table_cap;
    W4<T> *s_w = reinterpret_cast<W4<T> *>(smem);
    uint2 *s_nbr = reinterpret_cast<uint2 *>(s_w + cap);
    unsigned short *s_pix = reinterpret_cast<unsigned short *>(s_nbr + cap);
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_pix + cap); // [1 + G * STAGES] (cap is a multiple of 32)
    double *s_gsum = reinterpret_cast<double *>(s_bar + 1 + (size_t)G * Ring::STAGES + ((1 + G * Ring::STAGES) & 1)); // 16-byte aligned
    T *s_img = reinterpret_cast<T *>(s_gsum + (size_t)G * SB_FAST_MAXC);
    unsigned char *s_ring = reinterpret_cast<unsigned char *>(s_img + (size_t)G * wa.npix);

    const DevChain &ch = a.chains[a.src[list[0]].chain];
    const DevMono &mo = a.monos[ch.ops[0].iarg]; // host: the chain starts with the monotonic operator (fused pattern)
    // The CTA's operator table: three bulk asynchronous copies issued by one thread, complete on s_bar[0]; every warp waits for
    // them only right before its first sweep, i.e. the copy runs under the gradient gather of phase 1.
    const unsigned bar_tab = smem_u32(s_bar);
    if (threadIdx.x == 0) {
        mbar_init(bar_tab, 1);
        for (int i = 0; i < G * Ring::STAGES; ++i) mbar_init(bar_tab + 8u * (unsigned)(1 + i), 1);
        mbar_fence_init();
    }
    // is any source of the list still running?  (threads look at different entries; barrier + OR)
    int live_any = 0;
    for (int i = threadIdx.x; i < n_list; i += blockDim.x) live_any |= !a.done[a.src[list[i]].scene];
    if (!__syncthreads_or(live_any)) return; // every scene of this chain has stopped: nothing is copied
    if (wa.bulk_table) {
        if (threadIdx.x == 0) {
            const unsigned n = (unsigned)mo.w_cap;
            const unsigned char *src = static_cast<const unsigned char *>(mo.wtab);
            mbar_expect_tx(bar_tab, n * (unsigned)(sizeof(W4<T>) + sizeof(uint2) + sizeof(unsigned short)));
            bulk_g2s(smem_u32(s_w), src, n * (unsigned)sizeof(W4<T>), bar_tab);
            bulk_g2s(smem_u32(s_nbr), src + (size_t)n * sizeof(W4<T>), n * (unsigned)sizeof(uint2), bar_tab);
            bulk_g2s(smem_u32(s_pix), src + (size_t)n * (sizeof(W4<T>) + sizeof(uint2)), n * (unsigned)sizeof(unsigned short), bar_tab);
        }
    } else { // diagnostic: the same precomputed table copied by all threads
        const unsigned n = (unsigned)mo.w_cap;
        const uint4 *src = static_cast<const uint4 *>(mo.wtab);
        uint4 *dw = reinterpret_cast<uint4 *>(s_w), *dn = reinterpret_cast<uint4 *>(s_nbr), *dp = reinterpret_cast<uint4 *>(s_pix);
        const unsigned nw = n * (unsigned)sizeof(W4<T>) / 16u, nn2 = n * 8u / 16u, np = n * 2u / 16u;
        for (unsigned i = threadIdx.x; i < nw; i += blockDim.x) dw[i] = src[i];
        for (unsigned i = threadIdx.x; i < nn2; i += blockDim.x) dn[i] = src[nw + i];
        for (unsigned i = threadIdx.x; i < np; i += blockDim.x) dp[i] = src[nw + nn2 + i];
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tab) : "memory");
    }
    WarpTable<T> tab;
    tab.nbr = smem_u32(s_nbr + lane), tab.w = smem_u32(s_w + lane), tab.pix = smem_u32(s_pix + lane), tab.n_trips = mo.w_trips;
    // ring state of this warp, carried from source to source (every source leaves the ring drained)
    int iss_stage = 0, use_stage = 0, n_flight = 0;
    unsigned use_parity = 0u;
#pragma unroll 1
    for (;;) {
    int claim = 0;
    if (lane == 0) claim = atomicAdd(counter, 1);
    claim = __shfl_sync(0xffffffffu, claim, 0);
    if (claim >= n_list) break;
    const int k = list[claim];
    const DevSource &d = a.src[k];
    const int s = d.scene;
    if (a.done[s]) continue;
    const int it = a.it_ptr[s], C = a.C, n = d.By * d.Bx, Bx = d.Bx;
    const unsigned magic = 0xffffffffu / (unsigned)Bx + 1u; // p / Bx == umulhi(p, magic) for p, Bx < 65536
    T *zn = s_img + (size_t)wid * wa.npix;
    const unsigned zb = smem_u32(zn);
    T *mp = a.morph + d.morph_off, *mm = a.morph_m + d.morph_off, *mv = a.morph_v + d.morph_off, *mvh = a.morph_vhat + d.morph_off;
    XP<T> *xp = wa.xp + (size_t)(blockIdx.x * G + wid) * (32 * NPT); // this warp's padded, 16-byte aligned slot
    const bool upd = !d.morph_fixed;
    const T alpha = (T)d.morph_step;

    // ---- phase 1: gradient gather, spectrum-gradient partial sums, AMSGrad step -> x (shared + global), psi (global)
    double sedv[SB_FAST_MAXC];
#pragma unroll
    for (int c = 0; c < SB_FAST_MAXC; ++c) sedv[c] = c < C ? a.sed[(size_t)k * C + c] : 0.0;
    T gs[SB_FAST_MAXC];
#pragma unroll
    for (int c = 0; c < SB_FAST_MAXC; ++c) gs[c] = T(0);
    if (lane == 0) zn[n] = T(0); // the spare cell
    T pmax = T(0);
    constexpr int PB = 2; // pixels per trip: every load of the trip first
#pragma unroll 1
    for (int p0 = lane; p0 < n; p0 += PB * 32) {
        T mval[PB], m0[PB], v0[PB], vh0[PB], gv[PB][SB_FAST_MAXC];
        bool in[PB];
#pragma unroll
        for (int i = 0; i < PB; ++i) {
            const int p = p0 + i * 32;
            mval[i] = m0[i] = v0[i] = vh0[i] = T(0);
            in[i] = false;
#pragma unroll
            for (int c = 0; c < SB_FAST_MAXC; ++c) gv[i][c] = T(0);
            if (p < n) {
                const int by = (int)__umulhi((unsigned)p, magic), bx = p - by * Bx, y = d.oy + by, x = d.ox + bx;
                mval[i] = mp[p];
                if (upd) m0[i] = mm[p], v0[i] = mv[p], vh0[i] = mvh[p];
                in[i] = (unsigned)y < (unsigned)a.Ny && (unsigned)x < (unsigned)a.Nx;
                if (in[i]) {
                    if (a.n_obs == 1) {
                        grad_bands<T, SB_FAST_MAXC>(a, s, C, y, x, gv[i]);
                    } else {
                        grad_bands_multi<T, SB_FAST_MAXC>(a, s, C, y, x, gv[i]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < PB; ++i) {
            const int p = p0 + i * 32;
            if (p < n) {
                double gm = 0.0;
#pragma unroll
                for (int c = 0; c < SB_FAST_MAXC; ++c) {
                    if (c < C) {
                        gm += sedv[c] * (double)gv[i][c];
                        gs[c] += gv[i][c] * mval[i];
                    }
                }
                if (upd) {
                    T m_ = m0[i], v_ = v0[i], vh_ = vh0[i];
                    const T psi = amsgrad_t((T)gm, m_, v_, vh_, it, a.fs);
                    mm[p] = m_, mv[p] = v_, mvh[p] = vh_;
                    const T xn = mval[i] - alpha * m_ / psi;
                    zn[p] = xn; // first proximal argument: z0 - psi/max(psi) (z0 - x) = x exactly (z0 = x)
                    xp[p] = XP<T>{xn, psi};
                    pmax = psi > pmax ? psi : pmax;
                }
            }
        }
    }
    // spectrum gradient: shuffle reduction in double, parked in shared memory until the spectrum update at the end
    double *gsum = s_gsum + (size_t)wid * SB_FAST_MAXC;
#pragma unroll
    for (int c = 0; c < SB_FAST_MAXC; ++c) {
        const double t = c < C ? warp_sum_all((double)gs[c]) : 0.0;
        if (lane == 0) gsum[c] = t;
    }
    __syncwarp();

    if (upd) {
        const T psimax = warp_max_all_t<T>(pmax);
        const T fac = (alpha / psimax) / alpha; // gamma / alpha with gamma = alpha / max(psi), evaluated like update_extended
        const FusedChain fc = fused_chain_of(ch, d.By, d.Bx);
        const T hs = (T)(0.5 * fc.sym), om = (T)(1.0 - fc.sym), zero = (T)fc.zero, tiny = (T)fc.tiny, keep = T(1) - (T)fc.mono_grad;
        const T e2 = (T)(a.fs.e_rel * a.fs.e_rel);
        const int half = (n - 1) >> 1; // centre pixel (odd x odd box): its 180-degree partner is itself
        // the running iterate z (= x before the first projection): NPT registers per lane
        T zold[NPT];
#pragma unroll
        for (int i = 0; i < NPT; ++i) {
            const int p = lane + 32 * i;
            zold[i] = p < n ? zn[p] : T(0);
        }
        // ---- x / psi stream: chunk g of the stream (pass g / nch, chunk g % nch of the source's slot) lands in ring stage
        // g % STAGES; lane 0 keeps STAGES chunks in flight, speculating that the next pass happens (passes 0 .. prox_max - 2
        // read x / psi; what was requested but never consumed is drained before the warp leaves)
        const int prox_max = prox_max_iter_of(a, s);
        const int nch = (((n + 31) >> 5) + Ring::TRIPS - 1) / Ring::TRIPS; // chunks per pass
        const int g_end = nch * (prox_max - 1);                            // chunks of the whole stream
        const unsigned ring = smem_u32(s_ring + (size_t)wid * Ring::BYTES), bar_ring = bar_tab + 8u * (unsigned)(1 + wid * Ring::STAGES);
        // producer side (lane 0 issues; every lane keeps the same counters): stage, chunk within the pass, chunks left to request;
        // consumer side: stage, its phase parity, requested-but-unconsumed chunks (iss_stage, use_stage, use_parity, n_flight)
        int iss_chunk = 0, iss_left = g_end > 0 ? g_end : 0;
        auto issue = [&]() {
            if (lane == 0) {
                mbar_expect_tx(bar_ring + 8u * (unsigned)iss_stage, (unsigned)Ring::CHUNK);
                bulk_g2s(ring + (unsigned)(iss_stage * Ring::CHUNK), reinterpret_cast<const unsigned char *>(xp) + (size_t)iss_chunk * Ring::CHUNK,
                         (unsigned)Ring::CHUNK, bar_ring + 8u * (unsigned)iss_stage);
            }
            iss_stage = (iss_stage + 1) & (Ring::STAGES - 1);
            iss_chunk = iss_chunk + 1 == nch ? 0 : iss_chunk + 1;
            --iss_left, ++n_flight;
        };
        fence_proxy_async(); // this lane's st.global of x / psi -> visible to the bulk copies
        __syncwarp();
        if (!wa.use_ring) iss_left = 0;
        while (n_flight < Ring::STAGES && iss_left > 0) issue();
        mbar_wait(bar_tab, 0u); // the operator table has landed
        int nsub = 0;
#pragma unroll 1
        for (int sub = 0; sub < prox_max; ++sub) {
            warp_sweep<T>(zb, tab, keep);
            // ---- pass A: [symmetry] + positivity + centre floor, written back only when pixels are mixed; running maximum
            T mx = -INFINITY;
            if (fc.has_sym) {
#pragma unroll 1
                for (int p0 = lane; p0 <= half; p0 += 64) { // two pairs (p, n-1-p) per trip, their four loads first
                    T uu[2], vv[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int p = p0 + i * 32;
                        uu[i] = vv[i] = T(0);
                        if (p <= half) uu[i] = zn[p], vv[i] = zn[n - 1 - p];
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int p = p0 + i * 32;
                        if (p <= half) {
                            T u = hs * (uu[i] + vv[i]) + om * uu[i], v = hs * (vv[i] + uu[i]) + om * vv[i];
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
                __syncwarp();
            } else {
#pragma unroll
                for (int i = 0; i < NPT; ++i) {
                    const int p = lane + 32 * i;
                    if (p < n) {
                        T u = zn[p];
                        u = u < zero ? zero : u;
                        if (p == half) u = u < tiny ? tiny : u;
                        mx = u > mx ? u : mx;
                    }
                }
            }
            const T den = warp_max_all_t<T>(mx);
            const T inv = T(1) / den;
            // ---- pass B: normalise, convergence sums against the register-resident previous iterate, next argument
            const bool last = sub + 1 == prox_max;
            T dd = T(0), nn = T(0);
            constexpr int LB = Ring::TRIPS; // one ring chunk = the x / psi of LB trips
            static_assert(NPT % LB == 0, "NPT must be a multiple of the chunk length");
#pragma unroll
            for (int i0 = 0; i0 < NPT; i0 += LB) {
                if (32 * i0 < n) { // warp-uniform: chunk i0 / LB exists for this box
                    XP<T> q[LB];
#pragma unroll
                    for (int u = 0; u < LB; ++u) q[u] = XP<T>{T(0), T(0)};
                    if (!last && !wa.use_ring) {
#pragma unroll
                        for (int u = 0; u < LB; ++u) {
                            const int p = lane + 32 * (i0 + u);
                            if (p < n) q[u] = xp[p];
                        }
                    } else if (!last) {
                        mbar_wait(bar_ring + 8u * (unsigned)use_stage, use_parity);
                        const XP<T> *rq = reinterpret_cast<const XP<T> *>(s_ring + (size_t)wid * Ring::BYTES + (size_t)use_stage * Ring::CHUNK);
#pragma unroll
                        for (int u = 0; u < LB; ++u) q[u] = rq[32 * u + lane];
                    }
#pragma unroll
                    for (int u = 0; u < LB; ++u) {
                        const int i = i0 + u, p = lane + 32 * i;
                        if (p < n) {
                            T z = zn[p];
                            if (!fc.has_sym) {
                                z = z < zero ? zero : z;
                                if (p == half) z = z < tiny ? tiny : z;
                            }
                            T r;
                            if constexpr (sizeof(T) == 4)
                                r = z == den ? T(1) : z * inv;
                            else
                                r = z / den;
                            const T zo = zold[i], df = r - zo;
                            dd += df * df;
                            nn += zo * zo;
                            zold[i] = r;
                            if (!last) zn[p] = r - (fac * q[u].psi) * (r - q[u].x);
                        }
                    }
                    if (!last && wa.use_ring) { // the chunk's values have been consumed by every lane: refill its stage
                        __syncwarp();
                        use_stage = (use_stage + 1) & (Ring::STAGES - 1);
                        use_parity ^= use_stage == 0 ? 1u : 0u;
                        --n_flight;
                        if (iss_left > 0) issue();
                    }
                }
            }
            dd = warp_sum_all_t(dd);
            nn = warp_sum_all_t(nn);
            __syncwarp(); // zn is complete before the next sweep
            nsub = sub + 1;
            if (dd <= e2 * nn) break;
        }
        for (; n_flight > 0; --n_flight) { // requested for a pass that never came: let the copies land before the warp leaves
            mbar_wait(bar_ring + 8u * (unsigned)use_stage, use_parity);
            use_stage = (use_stage + 1) & (Ring::STAGES - 1);
            use_parity ^= use_stage == 0 ? 1u : 0u;
        }
        if (a.prox_hist && lane == 0) atomicAdd(a.prox_hist + min(nsub, 15), 1ull);
        // ---- the projected image -> morphology; a non-finite pixel poisons the sum
        T chk = T(0);
#pragma unroll
        for (int i = 0; i < NPT; ++i) {
            const int p = lane + 32 * i;
            if (p < n) {
                mp[p] = zold[i];
                chk += zold[i];
            }
        }
        chk = warp_sum_all_t(chk);
        if (lane == 0 && !isfinite((double)chk)) atomicExch(a.status + s, SB_ERR_NONFINITE);
    }
    if (lane == 0 && !d.sed_fixed) sed_update<T>(a, d, k, gsum, it);
    __syncwarp(); // the image and the spectrum sums of this source are dead: the next one may overwrite them
    }
    mbar_wait(bar_tab, 0u); // nobody leaves while the table copies are in flight
}

} // namespace sb
