// Fused spectral-convolution kernels of the fitting loop (sm_100a).
//
// The reference convolves the model cube with the per-band PSF difference kernel by padding, transforming with
// numpy.fft.rfftn, multiplying, transforming back and cropping (scarlet/renderer.py:215-259, scarlet/fft.py:200-273,
// 316-396), and autograd runs the same pipeline backwards for the gradient.  Here the 2-D transform is split into a
// row pass and a column pass, and every pass is fused with the work next to it, so that neither the zero-padded real
// grids nor the cropped-away part of the result ever touch HBM:
//
//   k_spec_render    render sed x morph (component.py:144-171, blend.py:200-244) for 2*npair frame rows in shared
//                    memory, real-to-complex row FFTs (two real rows packed as one complex transform), store the
//                    Ny x (Fx/2+1) half spectra X                                   [writes X]
//   k_spec_column    column FFT of the Ny non-zero rows (length Fy), multiply by K^ (or conj K^ for the adjoint,
//                    fft.py:316-331), inverse column FFT, keep rows [0,Ny)              [X in place, reads K^]
//   k_spec_residual  inverse row FFTs -> rendered model rows; r = w (m - d), chi^2 partial sums
//                    (observation.py:147-170, renderer.py:130-161); forward row FFTs of r   [X in place, reads d, w]
//   k_spec_column    (adjoint)
//   k_spec_grad      inverse row FFTs -> gradient of the loss wrt the model, rows [0,Ny) x [0,Nx)   [writes G]
//
//   k_spec_column_tma              the column pass with its tiles moved by the Tensor Memory Accelerator (float)
//   k_spec_column_fwd / _inv       the two halves of the column pass for observations on another pixel grid
//   k_resample_t1 / _lr / _q       ResolutionRenderer, aligned grids: separable Fourier resampling (renderer.py:262-547)
//   k_rot_partial / _residual / _adjoint   ResolutionRenderer, rotated grids: dense half-plane contraction (renderer.py:318-363)
//
// X: [S][C][Ny][Xp] complex, G: [S][C][Ny][Nx] real.  All transforms are unnormalised; 1/(Fy Fx) is folded into K^.
#pragma once
#include <cuda.h> // CUtensorMap

#include "common.cuh"
#include "fft_core.cuh"

namespace sb {

template <typename T> struct SpecObs {
    int C, H, W, chan_off, oy, ox; // data cube and its placement in the model frame
    int Fy, Fx, Fxc, Xp;           // grid, half-spectrum width Fx/2+1, row pitch of X and K^ (complex elements)
    int khat_shared;
    typename Cx<T>::type *X;          // [S][C][Ny][Xp]
    const typename Cx<T>::type *khat; // [S or 1][C][Fy][Xp]
    T *G;                             // [S][C][Ny][Nx]
    const T *data, *weights;          // [S][C][H][W]
    const typename Cx<T>::type *tw_x, *tw_y; // [R1][R2] tables exp(-2 pi i n2 k1 / L) for L = Fx and L = Fy
    // resampling observations (ResolutionRenderer, renderer.py:262-547): full-height spectra, shift matrices
    typename Cx<T>::type *P;          // [S][C][Fy][Xp]  K^ conj(M^)  /  h^2 K^ Q^
    typename Cx<T>::type *T1;         // [S][C][H][Xp]   Ey P  /  R Ex
    const typename Cx<T>::type *Ey;   // [H][Fy]   exp(-2 pi i f_ky ys_i), Nyquist real
    const typename Cx<T>::type *Ex;   // [W][Fxc]  exp(-2 pi i f_kx xs_j), Nyquist real
    T h2;                             // (pixel-scale ratio)^2
    // rotated resampling observations (renderer.py:318-363, 498-524): half-plane multipliers of the two-axis shifts
    const typename Cx<T>::type *RA;   // [H][Fy][Xp]  attached to low-resolution row i
    const typename Cx<T>::type *RB;   // [W][Fy][Xp]  attached to low-resolution column j
    T *Rres;                          // [S][C][H][W] weighted residual
    T *Rpart;                         // [S][C][n_chunk][H][W] partial sums of the render
    const int *cand_start, *cand;     // render kernel: per (scene, block of 2 npair rows) the sources whose boxes touch the rows
    int n_chunk, chunk;               // chunks of the flattened (ky, kx) index, entries per chunk (multiple of SB_ROT_SUB)
};

#define SB_SPEC_MAXCB 8 // bands per CTA of the row kernels (SpecArgs::cb <= this)

template <typename T> struct SpecArgs {
    SpecObs<T> ob;
    int Ny, Nx, Cm; // model frame (Cm = number of model channels)
    int npair;      // row pairs per CTA (row kernels)
    int cb;         // bands per CTA (row kernels): blockIdx.z selects bands [z*cb, z*cb+cb)
    const int *done;
    // render
    const DevSource *src;
    const int *scene_src_start;
    const double *sed;
    const T *morph, *pmorph, *smorph; // smorph: Fourier-shifted images of the shifting sources
    T *model_out; // optional [S][Cm][Ny][Nx]
    // residual
    double *partials; // [S][gridDim.x]
    T *rendered_out;  // optional [S][C][H][W]
    T *resid_out;     // optional [S][C][Ny][Nx]: w (rendered - data) in the model frame (psf_shift gradient)
    int conj;         // column kernel: multiply by conj(K^)
    unsigned magic_nx; // 2^32 / Nx + 1: idx / Nx == umulhi(idx, magic_nx)
};

// what the render kernel needs to know about one source whose box intersects the CTA's rows
template <typename T> struct __align__(16) SpecCand {
    int oy, ox, By, Bx;
    const T *mp; // morphology image (or first per-band plane of a point source)
    int plane;
    int sync_before; // 1: a barrier before this source is added (its columns overlap an earlier source of the same barrier group)
    T sed[SB_SPEC_MAXCB];
};

// ---- shared helpers of the row kernels -------------------------------------------------------------------------
// natural-order spectrum Z of a packed row pair (row y real part, row y+1 imaginary part) -> half spectra A, B
template <typename T, int R1, int R2>
__device__ __forceinline__ void split_and_store(const SpecArgs<T> &a, const typename Cx<T>::type *fbuf, int NB, int s, int y0,
                                                int c0, int Cb) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    const SpecObs<T> &ob = a.ob;
    const int Fxc = ob.Fxc, Fx = ob.Fx, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int f = threadIdx.x >> 5; f < NB; f += nw) { // one warp per transform: no per-element index arithmetic
        const int p = f / Cb, c = c0 + f - p * Cb, y = y0 + 2 * p;
        if (y >= a.Ny) continue;
        const C2 *z0 = fbuf + f * P::SF;
        C2 *row = ob.X + ((size_t)(s * ob.C + c) * a.Ny + y) * ob.Xp;
        const bool two = y + 1 < a.Ny;
        for (int k = lane; k < Fxc; k += 32) {
            const C2 z = z0[k], zc = z0[k ? Fx - k : 0];
            C2 A, B;
            A.x = T(0.5) * (z.x + zc.x), A.y = T(0.5) * (z.y - zc.y);
            B.x = T(0.5) * (z.y + zc.y), B.y = T(0.5) * (zc.x - z.x);
            row[k] = A;
            if (two) row[ob.Xp + k] = B;
        }
    }
}

// half spectra of rows y, y+1 -> natural-order full spectrum of the packed pair in fbuf
template <typename T, int R1, int R2>
__device__ __forceinline__ void load_and_merge(const SpecArgs<T> &a, typename Cx<T>::type *fbuf, int NB, int s, int y0, int c0,
                                               int Cb) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    const SpecObs<T> &ob = a.ob;
    constexpr int Fx = R1 * R2, Fxc = Fx / 2 + 1, TR = (Fxc + 31) / 32; // the row transform length is the template's
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int f = threadIdx.x >> 5; f < NB; f += nw) {
        const int p = f / Cb, c = c0 + f - p * Cb, y = y0 + 2 * p;
        C2 *z0 = fbuf + f * P::SF;
        const bool one = y < a.Ny, two = y + 1 < a.Ny;
        const C2 *row = ob.X + ((size_t)(s * ob.C + c) * a.Ny + (one ? y : 0)) * ob.Xp;
        // every load of the two half spectra first (2 TR independent requests per lane), then the merge: with the loop
        // rolled, each trip waited for its own two loads and a CTA spent most of its life in these round trips
        C2 A[TR], B[TR];
#pragma unroll
        for (int t = 0; t < TR; ++t) {
            const int k = lane + 32 * t;
            A[t] = B[t] = C2{T(0), T(0)};
            if (k < Fxc) {
                if (one) A[t] = row[k];
                if (two) B[t] = row[ob.Xp + k];
            }
        }
#pragma unroll
        for (int t = 0; t < TR; ++t) {
            const int k = lane + 32 * t;
            if (k < Fxc) {
                z0[k] = C2{A[t].x - B[t].y, A[t].y + B[t].x};
                if (k > 0 && 2 * k < Fx) z0[Fx - k] = C2{A[t].x + B[t].y, B[t].x - A[t].y};
            }
        }
    }
}

template <typename T, int R1, int R2>
__device__ __forceinline__ void stage_twiddles(typename Cx<T>::type *dst, const typename Cx<T>::type *src) {
    for (int i = threadIdx.x; i < R1 * R2; i += blockDim.x) dst[i] = src[i];
}

// inverse row transform of the merged spectra in fbuf: lane (f, n2) ends with a_[n1] = (row y, row y+1) at x = n1 R2 + n2
template <typename T, int R1, int R2>
__device__ __forceinline__ void rows_inverse(typename Cx<T>::type (&a_)[R1], typename Cx<T>::type *fbuf,
                                             const typename Cx<T>::type *tw, int NB) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    const int tid = threadIdx.x;
    C2 b_[R2];
    const int fb = tid / R1, k1 = tid - fb * R1;
    if (tid < NB * R1) sbfft::static_for<0, R2>([&](auto i) { b_[decltype(i)::value] = fbuf[fb * P::SF + k1 + R1 * decltype(i)::value]; });
    __syncthreads();
    if (tid < NB * R1) sbfft::inv_stage_b<R1, R2>(b_, k1, tw, fbuf + fb * P::SF);
    __syncthreads();
    const int fa = tid / R2, n2 = tid - fa * R2;
    if (tid < NB * R2) sbfft::inv_stage_a<R1, R2>(a_, n2, fbuf + fa * P::SF);
}

// forward row transform of a_ (lane (f, n2)); leaves the natural-order spectrum in fbuf
template <typename T, int R1, int R2>
__device__ __forceinline__ void rows_forward(typename Cx<T>::type (&a_)[R1], typename Cx<T>::type *fbuf,
                                             const typename Cx<T>::type *tw, int NB) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    const int tid = threadIdx.x;
    const int fa = tid / R2, n2 = tid - fa * R2;
    if (tid < NB * R2) sbfft::fwd_stage_a<R1, R2>(a_, n2, tw, fbuf + fa * P::SF);
    __syncthreads();
    C2 b_[R2];
    const int fb = tid / R1, k1 = tid - fb * R1;
    if (tid < NB * R1) sbfft::fwd_stage_b<R1, R2>(b_, k1, fbuf + fb * P::SF);
    __syncthreads();
    if (tid < NB * R1) sbfft::static_for<0, R2>([&](auto i) { fbuf[fb * P::SF + k1 + R1 * decltype(i)::value] = b_[decltype(i)::value]; });
    __syncthreads();
}

// ======================================================================================================
// render + forward row FFT
// ======================================================================================================
// E: pixels of one source a thread can hold ahead of the additions (2 where a typical box has more pixels in a row block than
// the CTA has threads; the one-pixel form is leaner where it suffices)
template <typename T, int R1, int R2, int E = 1> __global__ void __launch_bounds__(sizeof(T) == 4 ? 512 : 256) k_spec_render(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    extern __shared__ __align__(16) unsigned char smem[];
    const int s = blockIdx.y;
    if (a.done[s]) return;
    const SpecObs<T> &ob = a.ob;
    const int Co = ob.C, Nx = a.Nx, Ny = a.Ny, rows = 2 * a.npair, y0 = blockIdx.x * rows;
    const int c0 = blockIdx.z * a.cb, Cb = min(a.cb, Co - c0), NB = a.npair * Cb;
    const int tid = threadIdx.x, nt = blockDim.x;
    C2 *fbuf = reinterpret_cast<C2 *>(smem);
    C2 *tw = fbuf + (size_t)a.npair * a.cb * P::SF;
    T *tile = reinterpret_cast<T *>(tw + R1 * R2);
    SpecCand<T> *recs = reinterpret_cast<SpecCand<T> *>(
        (reinterpret_cast<uintptr_t>(tile + (size_t)a.cb * rows * Nx) + 15) & ~(uintptr_t)15);
    stage_twiddles<T, R1, R2>(tw, ob.tw_x);
    // Sources whose boxes intersect these rows: a list made once per plan (boxes do not move inside a plan; ob.cand_start /
    // ob.cand, scene order = the accumulation order of blend.py:17-27).  Thread j turns entry j into a compact record while
    // the others clear the tile; then the sources are added in list order, a thread per covered pixel -- the work is the
    // covered area, not (pixels of the rows) x (sources).
    const int blk = s * gridDim.x + blockIdx.x, l0 = ob.cand_start[blk], ncand = ob.cand_start[blk + 1] - l0;
    for (int j = tid; j < ncand; j += nt) {
        const int entry = ob.cand[l0 + j], k = entry & 0x7fffffff; // bit 31: barrier before this source
        const DevSource &d = a.src[k];
        SpecCand<T> rc;
        rc.oy = d.oy, rc.ox = d.ox, rc.By = d.By, rc.Bx = d.Bx;
        const int plane = d.kind == 0 ? 0 : d.By * d.Bx;
        rc.plane = plane, rc.sync_before = (int)((unsigned)entry >> 31);
        rc.mp = (d.kind == 0 ? (d.shifting ? a.smorph : a.morph) : a.pmorph + (size_t)(ob.chan_off + c0) * plane) + d.morph_off;
        const double *sed = a.sed + (size_t)k * a.Cm + ob.chan_off + c0;
#pragma unroll
        for (int c = 0; c < SB_SPEC_MAXCB; ++c) rc.sed[c] = c < Cb ? (T)sed[c] : T(0);
        recs[j] = rc;
    }
    for (int idx = tid; idx < Cb * rows * Nx; idx += nt) tile[idx] = T(0);
    __syncthreads();
    // Sources four at a time: every thread first requests its pixel of each of the four (one memory round trip for the
    // chunk), then the four are added in list order.  A barrier separates two sources only where the host found that their
    // columns overlap inside these rows (SpecCand::sync_before): sources side by side are added without waiting for each
    // other.  Boxes with more pixels in these rows than the CTA has threads, and point sources (one morphology plane per
    // band), load and add in one go when their turn comes.
#pragma unroll 1
    for (int i0 = 0; i0 < ncand; i0 += 4) {
        T v[4][E];
        int tp[4][E]; // tile offsets, -1: none
        bool plain[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < E; ++e) v[j][e] = T(0), tp[j][e] = -1;
            plain[j] = false;
            if (i0 + j < ncand) {
                const SpecCand<T> &rc = recs[i0 + j];
                const int ry0 = max(y0, rc.oy), ry1 = min(min(y0 + rows, Ny), rc.oy + rc.By); // rows of the box inside this CTA's rows
                const int bx0 = max(0, -rc.ox), bx1 = min(rc.Bx, Nx - rc.ox), wx = bx1 - bx0;  // columns of the box inside the frame
                const int npx = wx > 0 ? (ry1 - ry0) * wx : 0;
                plain[j] = npx > E * nt || rc.plane != 0;
                if (!plain[j]) {
                    const unsigned magic_w = wx > 0 ? 0xffffffffu / (unsigned)wx + 1u : 0u;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int idx = tid + e * nt;
                        if (idx < npx) {
                            const int rr = (int)__umulhi((unsigned)idx, magic_w), bx = bx0 + idx - rr * wx, y = ry0 + rr;
                            v[j][e] = rc.mp[(y - rc.oy) * rc.Bx + bx];
                            tp[j][e] = (y - y0) * Nx + rc.ox + bx;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j >= ncand) break; // CTA-uniform
            const SpecCand<T> &rc = recs[i0 + j];
            if (rc.sync_before) __syncthreads();
            if (!plain[j]) {
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (tp[j][e] >= 0) {
                        T *t = tile + tp[j][e];
#pragma unroll
                        for (int c = 0; c < SB_SPEC_MAXCB; ++c)
                            if (c < Cb) t[(size_t)c * rows * Nx] += rc.sed[c] * v[j][e];
                    }
                continue;
            }
            const int ry0 = max(y0, rc.oy), ry1 = min(min(y0 + rows, Ny), rc.oy + rc.By);
            const int bx0 = max(0, -rc.ox), bx1 = min(rc.Bx, Nx - rc.ox), wx = bx1 - bx0;
            const int npx = wx > 0 ? (ry1 - ry0) * wx : 0;
            const unsigned magic_w = wx > 0 ? 0xffffffffu / (unsigned)wx + 1u : 0u;
            const int plane = rc.plane;
            for (int idx = tid; idx < npx; idx += nt) {
                const int rr = (int)__umulhi((unsigned)idx, magic_w), bx = bx0 + idx - rr * wx, y = ry0 + rr;
                const T *pm = rc.mp + (y - rc.oy) * rc.Bx + bx;
                T *t = tile + (size_t)(y - y0) * Nx + rc.ox + bx;
                if (plane == 0) {
                    const T v1 = pm[0];
#pragma unroll
                    for (int c = 0; c < SB_SPEC_MAXCB; ++c)
                        if (c < Cb) t[(size_t)c * rows * Nx] += rc.sed[c] * v1;
                } else {
#pragma unroll
                    for (int c = 0; c < SB_SPEC_MAXCB; ++c)
                        if (c < Cb) t[(size_t)c * rows * Nx] += rc.sed[c] * pm[c * plane];
                }
            }
        }
    }
    __syncthreads(); // the tile is complete
    if (a.model_out) {
        for (int idx = tid; idx < Cb * rows * Nx; idx += nt) {
            const int cr = (int)__umulhi((unsigned)idx, a.magic_nx), x = idx - cr * Nx, c = cr / rows, y = y0 + cr - c * rows;
            if (y < Ny) a.model_out[(((size_t)s * a.Cm + ob.chan_off + c0 + c) * Ny + y) * Nx + x] = tile[idx];
        }
    }
    C2 a_[R1];
    {
        const int f = tid / R2, n2 = tid - f * R2;
        if (tid < NB * R2) {
            const int p = f / Cb, c = f - p * Cb;
            const T *t0 = tile + ((size_t)c * rows + 2 * p) * Nx, *t1 = t0 + Nx;
            sbfft::static_for<0, R1>([&](auto i) {
                const int n = decltype(i)::value * R2 + n2;
                a_[decltype(i)::value] = n < Nx ? C2{t0[n], t1[n]} : C2{T(0), T(0)};
            });
        }
    }
    rows_forward<T, R1, R2>(a_, fbuf, tw, NB);
    split_and_store<T, R1, R2>(a, fbuf, NB, s, y0, c0, Cb);
}

// ======================================================================================================
// inverse row FFT -> residual + loss -> forward row FFT
// ======================================================================================================
// WITH_RESID: also write the residual in the model frame (SpecArgs::resid_out; psf_shift plans only -- a separate instantiation
// keeps the extra stores and their predicates out of the hot kernel)
template <typename T, int R1, int R2, bool WITH_RESID = false>
__global__ void __launch_bounds__(sizeof(T) == 4 ? 512 : 256, 2) k_spec_residual(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[40];
    const int s = blockIdx.y;
    if (a.done[s]) return;
    const SpecObs<T> &ob = a.ob;
    const int Co = ob.C, Nx = a.Nx, Ny = a.Ny, rows = 2 * a.npair, y0 = blockIdx.x * rows;
    const int c0 = blockIdx.z * a.cb, Cb = min(a.cb, Co - c0), NB = a.npair * Cb;
    const int tid = threadIdx.x;
    C2 *fbuf = reinterpret_cast<C2 *>(smem);
    C2 *tw = fbuf + NB * P::SF;
    stage_twiddles<T, R1, R2>(tw, ob.tw_x);
    // data and weights are needed only after the inverse transform: pull their lines into L2 now (no registers held)
    {
        const int lines = (Nx * (int)sizeof(T) + 127) / 128; // 128-byte lines per row
        for (int idx = tid; idx < Cb * rows * lines; idx += blockDim.x) {
            const int l = idx % lines, rc = idx / lines, r = rc % rows, c = c0 + rc / rows;
            const int y = y0 + r, dy = y - ob.oy, dx = l * (128 / (int)sizeof(T)) - ob.ox;
            if (y < Ny && (unsigned)dy < (unsigned)ob.H && (unsigned)dx < (unsigned)ob.W) {
                const size_t off = ((size_t)(s * Co + c) * ob.H + dy) * ob.W + dx;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ob.data + off));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ob.weights + off));
            }
        }
    }
    load_and_merge<T, R1, R2>(a, fbuf, NB, s, y0, c0, Cb);
    __syncthreads();
    C2 a_[R1];
    rows_inverse<T, R1, R2>(a_, fbuf, tw, NB);
    T part_t = T(0); // <= 2 R1 terms per lane in the working precision; lanes are then summed in double
    {
        const int f = tid / R2, n2 = tid - f * R2;
        if (tid < NB * R2) {
            const int p = f / Cb, c = c0 + f - p * Cb, y = y0 + 2 * p;
            const int dy0 = y - ob.oy, dy1 = dy0 + 1;
            const bool row0 = y < Ny && (unsigned)dy0 < (unsigned)ob.H, row1 = y + 1 < Ny && (unsigned)dy1 < (unsigned)ob.H;
            const size_t base0 = ((size_t)(s * Co + c) * ob.H + dy0) * ob.W, base1 = base0 + ob.W;
            // Four columns at a time: their (up to) sixteen data / weight values are requested first, then consumed.  Element
            // by element, every pair of loads sat behind the branch (and the optional rendered_out store) of the previous
            // element and a lane paid 2 R1 dependent round trips to L2.
            constexpr int CH = 4;
            sbfft::static_for<0, (R1 + CH - 1) / CH>([&](auto gi) {
                constexpr int g0 = decltype(gi)::value * CH;
                T wv[CH][2], dv[CH][2];
                bool ok[CH][2];
                sbfft::static_for<0, CH>([&](auto ui) {
                    constexpr int u = decltype(ui)::value, n1 = g0 + u;
                    if (n1 < R1) {
                        const int x = n1 * R2 + n2, dx = x - ob.ox;
                        const bool col = x < Nx && (unsigned)dx < (unsigned)ob.W;
                        ok[u][0] = col && row0, ok[u][1] = col && row1;
                        wv[u][0] = dv[u][0] = wv[u][1] = dv[u][1] = T(0);
                        if (ok[u][0]) wv[u][0] = __ldg(ob.weights + base0 + dx), dv[u][0] = __ldg(ob.data + base0 + dx);
                        if (ok[u][1]) wv[u][1] = __ldg(ob.weights + base1 + dx), dv[u][1] = __ldg(ob.data + base1 + dx);
                    }
                });
                sbfft::static_for<0, CH>([&](auto ui) {
                    constexpr int u = decltype(ui)::value, n1 = g0 + u;
                    if (n1 < R1) {
                        const int dx = n1 * R2 + n2 - ob.ox;
                        T r0 = T(0), r1 = T(0);
                        if (ok[u][0]) {
                            const T m = a_[n1 < R1 ? n1 : 0].x, diff = m - dv[u][0];
                            r0 = wv[u][0] * diff;
                            part_t += r0 * diff;
                            if (a.rendered_out) a.rendered_out[base0 + dx] = m;
                        }
                        if (ok[u][1]) {
                            const T m = a_[n1 < R1 ? n1 : 0].y, diff = m - dv[u][1];
                            r1 = wv[u][1] * diff;
                            part_t += r1 * diff;
                            if (a.rendered_out) a.rendered_out[base1 + dx] = m;
                        }
                        a_[n1 < R1 ? n1 : 0] = C2{r0, r1};
                        if constexpr (WITH_RESID) {
                            const int x = n1 * R2 + n2;
                            if (x < Nx && y < Ny) {
                                T *ro = a.resid_out + ((size_t)(s * Co + c) * Ny + y) * Nx + x;
                                ro[0] = r0;
                                if (y + 1 < Ny) ro[Nx] = r1;
                            }
                        }
                    }
                });
            });
        }
    }
    __syncthreads(); // every lane has read its inverse-transform output before fbuf is reused
    rows_forward<T, R1, R2>(a_, fbuf, tw, NB);
    split_and_store<T, R1, R2>(a, fbuf, NB, s, y0, c0, Cb);
    const double part = block_sum((double)part_t, red);
    if (tid == 0) a.partials[((size_t)s * gridDim.z + blockIdx.z) * gridDim.x + blockIdx.x] = part;
}

// ======================================================================================================
// inverse row FFT -> gradient wrt the model
// ======================================================================================================
template <typename T, int R1, int R2> __global__ void __launch_bounds__(sizeof(T) == 4 ? 512 : 256) k_spec_grad(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    extern __shared__ __align__(16) unsigned char smem[];
    const int s = blockIdx.y;
    if (a.done[s]) return;
    const SpecObs<T> &ob = a.ob;
    const int Co = ob.C, Nx = a.Nx, Ny = a.Ny, rows = 2 * a.npair, y0 = blockIdx.x * rows;
    const int c0 = blockIdx.z * a.cb, Cb = min(a.cb, Co - c0), NB = a.npair * Cb;
    const int tid = threadIdx.x;
    C2 *fbuf = reinterpret_cast<C2 *>(smem);
    C2 *tw = fbuf + NB * P::SF;
    stage_twiddles<T, R1, R2>(tw, ob.tw_x);
    load_and_merge<T, R1, R2>(a, fbuf, NB, s, y0, c0, Cb);
    __syncthreads();
    C2 a_[R1];
    rows_inverse<T, R1, R2>(a_, fbuf, tw, NB);
    const int f = tid / R2, n2 = tid - f * R2;
    if (tid < NB * R2) {
        const int p = f / Cb, c = c0 + f - p * Cb, y = y0 + 2 * p;
        if (y < Ny) {
            T *g0 = ob.G + ((size_t)(s * Co + c) * Ny + y) * Nx;
            const bool row1 = y + 1 < Ny;
            sbfft::static_for<0, R1>([&](auto i) {
                constexpr int n1 = decltype(i)::value;
                const int x = n1 * R2 + n2;
                if (x < Nx) {
                    g0[x] = a_[n1].x;
                    if (row1) g0[Nx + x] = a_[n1].y;
                }
            });
        }
    }
}

// ======================================================================================================
// column pass: forward column FFT, x K^ (or conj K^), inverse column FFT; NB adjacent columns per CTA
// ======================================================================================================
template <typename T, int R1, int R2, int NB> __global__ void __launch_bounds__(NB *sbfft::Plan2<R1, R2>::RMAX, (sizeof(T) == 4 && sbfft::Plan2<R1, R2>::RMAX <= 20) ? 3 : 1) k_spec_column(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    extern __shared__ __align__(16) unsigned char smem[];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.y, s = img / ob.C;
    if (a.done[s]) return;
    const int Ny = a.Ny, tid = threadIdx.x;
    C2 *fbuf = reinterpret_cast<C2 *>(smem);
    C2 *tw = fbuf + NB * P::SF;
    stage_twiddles<T, R1, R2>(tw, ob.tw_y);
    const int f = tid % NB, j = tid / NB, kx = blockIdx.x * NB + f;
    const bool col = kx < ob.Fxc;
    C2 *sm = fbuf + f * P::SF;
    C2 *X = ob.X + (size_t)img * Ny * ob.Xp + kx;
    const C2 *K = ob.khat + (size_t)(ob.khat_shared ? img - s * ob.C : img) * ob.Fy * ob.Xp + kx;
    // the K^ column of this lane is requested first: its R2 loads are in flight during the X loads and the whole forward transform
    C2 k_[R2];
    if (j < R1 && col)
        sbfft::static_for<0, R2>([&](auto i) { k_[decltype(i)::value] = K[(size_t)(j + R1 * decltype(i)::value) * ob.Xp]; });
    __syncthreads();
    C2 a_[R1];
    if (j < R2) {
        sbfft::static_for<0, R1>([&](auto i) {
            const int n = decltype(i)::value * R2 + j;
            a_[decltype(i)::value] = (col && n < Ny) ? X[(size_t)n * ob.Xp] : C2{T(0), T(0)};
        });
        sbfft::fwd_stage_a<R1, R2>(a_, j, tw, sm);
    }
    __syncthreads();
    C2 b_[R2];
    if (j < R1) {
        sbfft::fwd_stage_b<R1, R2>(b_, j, sm);
        if (col) {
            if (a.conj)
                sbfft::static_for<0, R2>([&](auto i) {
                    constexpr int k2 = decltype(i)::value;
                    b_[k2] = sbfft::cmul_conj(b_[k2], k_[k2]);
                });
            else
                sbfft::static_for<0, R2>([&](auto i) {
                    constexpr int k2 = decltype(i)::value;
                    b_[k2] = sbfft::cmul(b_[k2], k_[k2]);
                });
        }
    }
    __syncthreads();
    if (j < R1) sbfft::inv_stage_b<R1, R2>(b_, j, tw, sm);
    __syncthreads();
    if (j < R2) {
        sbfft::inv_stage_a<R1, R2>(a_, j, sm);
        if (col)
            sbfft::static_for<0, R1>([&](auto i) {
                const int n = decltype(i)::value * R2 + j;
                if (n < Ny) X[(size_t)n * ob.Xp] = a_[decltype(i)::value];
            });
    }
}

// ---- the same column pass with the tiles moved by the Tensor Memory Accelerator ------------------------------------------
// One thread issues three tensor copies (cp.async.bulk.tensor.2d, SASS UTMALDG): the [Ny x NB] tile of X and the [Fy x NB]
// tile of K^ (two boxes of Fy/2 rows: a box dimension is limited to 256) land in shared memory and complete on an mbarrier;
// the transformed tile goes back with one tensor store (UTMASTG).  No thread computes a global address or issues a global
// load / store; columns beyond the pitch are zero-filled on load and clipped on store by the copy engine.  The exchange
// buffer of the two-stage FFT aliases the tiles (they are dead once every thread holds its elements in registers), so a CTA
// needs (Ny + Fy) NB sizeof(complex) bytes: three CTAs per SM at 288 x 288, whose copies overlap each other's arithmetic.
namespace tma {
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wait(unsigned bar, unsigned parity) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "W_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra D_%=;\n"
                 "bra W_%=;\n"
                 "D_%=:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void load_2d(unsigned dst, const CUtensorMap *map, int x, int y, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void store_2d(const CUtensorMap *map, int x, int y, unsigned src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
} // namespace tma

template <typename T, int R1, int R2, int NB>
__global__ void __launch_bounds__(NB *sbfft::Plan2<R1, R2>::RMAX, (sizeof(T) == 4 && sbfft::Plan2<R1, R2>::RMAX <= 20) ? 3 : 1)
    k_spec_column_tma(const SpecArgs<T> a, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmK) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.y, s = img / ob.C;
    if (a.done[s]) return;
    const int Ny = a.Ny, Fy = ob.Fy, tid = threadIdx.x;
    // (aligned by an offset, not through an integer cast: the compiler keeps the shared address space and emits LDS / STS)
    unsigned char *smem = smem_raw + ((128u - (tma::smem_addr(smem_raw) & 127u)) & 127u);
    C2 *tileX = reinterpret_cast<C2 *>(smem);                 // [Ny][NB]
    C2 *tileK = tileX + (size_t)Ny * NB;                      // [Fy][NB]
    C2 *tw = tileK + (size_t)Fy * NB;                         // [R1 R2]
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(tw + R1 * R2);
    C2 *fbuf = tileX;                                         // exchange buffer: aliases the tiles (see above)
    const unsigned bar_a = tma::smem_addr(bar);
    const int kx0 = blockIdx.x * NB, kimg = ob.khat_shared ? img - s * ob.C : img;
    if (tid == 0) {
        tma::mbar_init(bar_a, 1);
        tma::fence_barrier_init();
        tma::expect_tx(bar_a, (unsigned)((Ny + Fy) * NB * sizeof(C2)));
        tma::load_2d(tma::smem_addr(tileX), &tmX, kx0, img * Ny, bar_a);
        tma::load_2d(tma::smem_addr(tileK), &tmK, kx0, kimg * Fy, bar_a);
        tma::load_2d(tma::smem_addr(tileK + (size_t)(Fy / 2) * NB), &tmK, kx0, kimg * Fy + Fy / 2, bar_a);
    }
    stage_twiddles<T, R1, R2>(tw, ob.tw_y);
    __syncthreads(); // the barrier is initialised for everybody, the twiddles are staged
    tma::wait(bar_a, 0u);
    const int f = tid % NB, j = tid / NB;
    C2 *sm = fbuf + f * P::SF;
    C2 k_[R2], a_[R1];
    if (j < R1) sbfft::static_for<0, R2>([&](auto i) { k_[decltype(i)::value] = tileK[(size_t)(j + R1 * decltype(i)::value) * NB + f]; });
    if (j < R2)
        sbfft::static_for<0, R1>([&](auto i) {
            const int n = decltype(i)::value * R2 + j;
            a_[decltype(i)::value] = n < Ny ? tileX[(size_t)n * NB + f] : C2{T(0), T(0)};
        });
    __syncthreads(); // every element sits in a register: the tiles may be overwritten by the exchange buffer
    if (j < R2) sbfft::fwd_stage_a<R1, R2>(a_, j, tw, sm);
    __syncthreads();
    C2 b_[R2];
    if (j < R1) {
        sbfft::fwd_stage_b<R1, R2>(b_, j, sm);
        if (a.conj)
            sbfft::static_for<0, R2>([&](auto i) {
                constexpr int k2 = decltype(i)::value;
                b_[k2] = sbfft::cmul_conj(b_[k2], k_[k2]);
            });
        else
            sbfft::static_for<0, R2>([&](auto i) {
                constexpr int k2 = decltype(i)::value;
                b_[k2] = sbfft::cmul(b_[k2], k_[k2]);
            });
    }
    __syncthreads();
    if (j < R1) sbfft::inv_stage_b<R1, R2>(b_, j, tw, sm);
    __syncthreads();
    if (j < R2) sbfft::inv_stage_a<R1, R2>(a_, j, sm);
    __syncthreads(); // the exchange buffer is dead: the result tile takes its place
    if (j < R2)
        sbfft::static_for<0, R1>([&](auto i) {
            const int n = decltype(i)::value * R2 + j;
            if (n < Ny) tileX[(size_t)n * NB + f] = a_[decltype(i)::value];
        });
    tma::fence_proxy_async_smem(); // st.shared -> visible to the copy engine
    __syncthreads();
    if (tid == 0) {
        tma::store_2d(&tmX, kx0, img * Ny, tma::smem_addr(tileX));
        tma::commit_and_wait_read(); // shared memory stays alive until the engine has read the tile
    }
}

// ======================================================================================================
// Resampling observation (ResolutionRenderer, scarlet/renderer.py:262-547).  With Parseval's theorem the reference's
// "Fourier-shift the kernel to every low-resolution row, Fourier-shift the model to every low-resolution column,
// multiply and sum" (its _resconv_op matrix product, renderer.py:478-547) is
//     LR[c,i,j] = h^2 sum_{ky,kx} Ey[i,ky] Ex[j,kx] K^[c,ky,kx] conj(M^[c,ky,kx])          (K^ carries 1/(Fy Fx))
// and its adjoint  G = h^2 IDFT2( K^ (Ey^T R Ex) ).  M^ comes from k_spec_render + k_spec_column_fwd, G leaves through
// k_spec_column_inv + k_spec_grad; in between sit three small dense contractions over ky, kx and (i, j).
// ======================================================================================================
// forward column FFT of the Ny non-zero rows; P = K^ conj(M^) for all Fy rows
template <typename T, int R1, int R2, int NB> __global__ void __launch_bounds__(NB *sbfft::Plan2<R1, R2>::RMAX) k_spec_column_fwd(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    extern __shared__ __align__(16) unsigned char smem[];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.y, s = img / ob.C;
    if (a.done[s]) return;
    const int Ny = a.Ny, tid = threadIdx.x;
    C2 *fbuf = reinterpret_cast<C2 *>(smem);
    C2 *tw = fbuf + NB * P::SF;
    stage_twiddles<T, R1, R2>(tw, ob.tw_y);
    const int f = tid % NB, j = tid / NB, kx = blockIdx.x * NB + f;
    const bool col = kx < ob.Fxc;
    C2 *sm = fbuf + f * P::SF;
    const C2 *X = ob.X + (size_t)img * Ny * ob.Xp + kx;
    __syncthreads();
    C2 a_[R1];
    if (j < R2) {
        sbfft::static_for<0, R1>([&](auto i) {
            const int n = decltype(i)::value * R2 + j;
            a_[decltype(i)::value] = (col && n < Ny) ? X[(size_t)n * ob.Xp] : C2{T(0), T(0)};
        });
        sbfft::fwd_stage_a<R1, R2>(a_, j, tw, sm);
    }
    __syncthreads();
    if (j < R1) {
        C2 b_[R2];
        sbfft::fwd_stage_b<R1, R2>(b_, j, sm);
        if (col) {
            const C2 *K = ob.khat + (size_t)(ob.khat_shared ? img - s * ob.C : img) * ob.Fy * ob.Xp + kx;
            C2 *Pout = ob.P + (size_t)img * ob.Fy * ob.Xp + kx;
            sbfft::static_for<0, R2>([&](auto i) {
                constexpr int k2 = decltype(i)::value;
                const size_t ky = j + R1 * k2;
                Pout[ky * ob.Xp] = sbfft::cmul_conj(K[ky * ob.Xp], b_[k2]); // K^ conj(M^)
            });
        }
    }
}

// inverse column FFT of full-height spectra P -> rows [0,Ny) of X
template <typename T, int R1, int R2, int NB> __global__ void __launch_bounds__(NB *sbfft::Plan2<R1, R2>::RMAX) k_spec_column_inv(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    typedef sbfft::Plan2<R1, R2> P;
    extern __shared__ __align__(16) unsigned char smem[];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.y, s = img / ob.C;
    if (a.done[s]) return;
    const int Ny = a.Ny, tid = threadIdx.x;
    C2 *fbuf = reinterpret_cast<C2 *>(smem);
    C2 *tw = fbuf + NB * P::SF;
    stage_twiddles<T, R1, R2>(tw, ob.tw_y);
    const int f = tid % NB, j = tid / NB, kx = blockIdx.x * NB + f;
    const bool col = kx < ob.Fxc;
    C2 *sm = fbuf + f * P::SF;
    __syncthreads();
    if (j < R1) {
        C2 b_[R2];
        const C2 *Pin = ob.P + (size_t)img * ob.Fy * ob.Xp + kx;
        sbfft::static_for<0, R2>([&](auto i) {
            constexpr int k2 = decltype(i)::value;
            b_[k2] = col ? Pin[(size_t)(j + R1 * k2) * ob.Xp] : C2{T(0), T(0)};
        });
        sbfft::inv_stage_b<R1, R2>(b_, j, tw, sm);
    }
    __syncthreads();
    if (j < R2) {
        C2 a_[R1];
        sbfft::inv_stage_a<R1, R2>(a_, j, sm);
        C2 *X = ob.X + (size_t)img * Ny * ob.Xp + kx;
        if (col)
            sbfft::static_for<0, R1>([&](auto i) {
                const int n = decltype(i)::value * R2 + j;
                if (n < Ny) X[(size_t)n * ob.Xp] = a_[decltype(i)::value];
            });
    }
}

// T1[img][i][kx] = sum_ky Ey[i][ky] P[img][ky][kx];  grid (ceil(Fxc/128), ceil(H/16), S*C), 128 threads: one kx per thread,
// 16 low-resolution rows per CTA whose Ey rows are staged in shared memory (P is read ceil(H/16) times in total)
// four consecutive reals by 128-bit shared / global loads
template <typename T> struct RotVec;
template <> struct RotVec<float> {
    static __device__ __forceinline__ void load4(const float *p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    }
};
template <> struct RotVec<double> {
    static __device__ __forceinline__ void load4(const double *p, double (&v)[4]) {
        const double2 t0 = *reinterpret_cast<const double2 *>(p), t1 = *reinterpret_cast<const double2 *>(p + 2);
        v[0] = t0.x, v[1] = t0.y, v[2] = t1.x, v[3] = t1.y;
    }
};
#define SB_RS_ROWS 16
template <typename T> __global__ void __launch_bounds__(128) k_resample_t1(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    extern __shared__ __align__(16) unsigned char smem[];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.z, s = img / ob.C;
    if (a.done[s]) return;
    const int kx = blockIdx.x * 128 + threadIdx.x, i0 = blockIdx.y * SB_RS_ROWS, ni = min(SB_RS_ROWS, ob.H - i0);
    C2 *ey = reinterpret_cast<C2 *>(smem); // [ky][SB_RS_ROWS]
    for (int idx = threadIdx.x; idx < ob.Fy * SB_RS_ROWS; idx += blockDim.x) {
        const int ky = idx / SB_RS_ROWS, q = idx - ky * SB_RS_ROWS;
        ey[idx] = q < ni ? ob.Ey[(size_t)(i0 + q) * ob.Fy + ky] : C2{T(0), T(0)};
    }
    __syncthreads();
    if (kx >= ob.Fxc) return;
    C2 acc[SB_RS_ROWS];
#pragma unroll
    for (int q = 0; q < SB_RS_ROWS; ++q) acc[q] = C2{T(0), T(0)};
    const C2 *Pc = ob.P + (size_t)img * ob.Fy * ob.Xp + kx;
    // four spectrum rows per trip, requested before they are used (rolled, every trip waited for its own L2 round trip);
    // the shift rows are read two complex entries per shared load
    constexpr int UN = 4;
#pragma unroll 1
    for (int ky = 0; ky < ob.Fy; ky += UN) {
        C2 p[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) p[u] = ky + u < ob.Fy ? Pc[(size_t)(ky + u) * ob.Xp] : C2{T(0), T(0)};
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (ky + u < ob.Fy) { // CTA-uniform
                const T *e = reinterpret_cast<const T *>(ey + (ky + u) * SB_RS_ROWS);
#pragma unroll
                for (int q = 0; q < SB_RS_ROWS; q += 2) {
                    T ev[4];
                    RotVec<T>::load4(e + 2 * q, ev);
                    acc[q].x += ev[0] * p[u].x - ev[1] * p[u].y;
                    acc[q].y += ev[0] * p[u].y + ev[1] * p[u].x;
                    acc[q + 1].x += ev[2] * p[u].x - ev[3] * p[u].y;
                    acc[q + 1].y += ev[2] * p[u].y + ev[3] * p[u].x;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < SB_RS_ROWS; ++q)
        if (q < ni) ob.T1[((size_t)img * ob.H + i0 + q) * ob.Xp + kx] = acc[q];
}

// one CTA per (scene, band): LR = h^2 sum_kx c_kx Re(Ex T1), residual r = w (LR - d), chi^2 partial, then U = r Ex -> T1
template <typename T> __global__ void __launch_bounds__(256) k_resample_lr(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ double red[40];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.x, s = img / ob.C;
    if (a.done[s]) return;
    const int H = ob.H, W = ob.W, Fxc = ob.Fxc, tid = threadIdx.x, nt = blockDim.x;
    // shared memory: the row-resampled spectra of this image [H][Fxc], the column shift table [W][Fxc], the residual [H][W]
    C2 *sT = reinterpret_cast<C2 *>(smem);
    C2 *sE = sT + (size_t)H * Fxc;
    T *r = reinterpret_cast<T *>(sE + (size_t)W * Fxc);
    C2 *T1 = ob.T1 + (size_t)img * H * ob.Xp;
    for (int idx = tid; idx < H * Fxc; idx += nt) {
        const int i = idx / Fxc, kx = idx - i * Fxc;
        sT[idx] = T1[(size_t)i * ob.Xp + kx];
    }
    for (int idx = tid; idx < W * Fxc; idx += nt) sE[idx] = ob.Ex[idx];
    __syncthreads();
    const bool even = (ob.Fx & 1) == 0;
    double part = 0.0;
    for (int idx = tid; idx < H * W; idx += nt) {
        const int i = idx / W, j = idx - i * W;
        const C2 *t = sT + (size_t)i * Fxc, *e = sE + (size_t)j * Fxc;
        // sum_kx c_kx Re(e t) with c = 1, 2, ..., 2, (1): twice the plain sum minus the end terms
        T acc0 = T(0), acc1 = T(0);
        int kx = 0;
        for (; kx + 1 < Fxc; kx += 2) {
            acc0 += e[kx].x * t[kx].x - e[kx].y * t[kx].y;
            acc1 += e[kx + 1].x * t[kx + 1].x - e[kx + 1].y * t[kx + 1].y;
        }
        if (kx < Fxc) acc0 += e[kx].x * t[kx].x - e[kx].y * t[kx].y;
        const T first = e[0].x * t[0].x - e[0].y * t[0].y;
        const T last = even ? e[Fxc - 1].x * t[Fxc - 1].x - e[Fxc - 1].y * t[Fxc - 1].y : T(0);
        const T m = ob.h2 * (T(2) * (acc0 + acc1) - first - last);
        const size_t di = (size_t)img * H * W + idx;
        const T w = ob.weights[di], diff = m - ob.data[di];
        r[idx] = w * diff;
        part += (double)w * (double)diff * (double)diff;
        if (a.rendered_out) a.rendered_out[di] = m;
    }
    __syncthreads();
    for (int idx = tid; idx < H * Fxc; idx += nt) {
        const int i = idx / Fxc, kx = idx - i * Fxc;
        C2 acc = C2{T(0), T(0)};
        for (int j = 0; j < W; ++j) {
            const C2 e = sE[(size_t)j * Fxc + kx];
            const T rv = r[i * W + j];
            acc.x += rv * e.x, acc.y += rv * e.y;
        }
        T1[(size_t)i * ob.Xp + kx] = acc;
    }
    part = block_sum(part, red);
    if (tid == 0) a.partials[img] = part;
}

// P[img][ky][kx] = h^2 K^[ky][kx] sum_i Ey[i][ky] U[img][i][kx];  grid (ceil(Fxc/128), ceil(Fy/64), S*C), 128 threads: one kx
// per thread, 64 ky per CTA.  The U column of a thread is held in registers 32 rows at a time (a low-resolution cube of up to
// 32 rows goes in one pass: P is written once, never read back), Ey in shared memory, four ky per trip: two 128-bit shared
// loads serve four outputs (eight independent accumulation chains) and the four K^ entries are requested ahead of the products.
#define SB_RS_KY 64
#define SB_RS_QROWS 32
template <typename T> __global__ void __launch_bounds__(128) k_resample_q(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    extern __shared__ __align__(16) unsigned char smem[];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.z, s = img / ob.C;
    if (a.done[s]) return;
    const int kx = blockIdx.x * 128 + threadIdx.x, ky0 = blockIdx.y * SB_RS_KY, nky = min(SB_RS_KY, ob.Fy - ky0);
    C2 *ey = reinterpret_cast<C2 *>(smem); // [H][SB_RS_KY]
    for (int idx = threadIdx.x; idx < ob.H * SB_RS_KY; idx += blockDim.x) {
        const int i = idx / SB_RS_KY, q = idx - i * SB_RS_KY;
        ey[idx] = q < nky ? ob.Ey[(size_t)i * ob.Fy + ky0 + q] : C2{T(0), T(0)};
    }
    __syncthreads();
    if (kx >= ob.Fxc) return;
    const C2 *U = ob.T1 + (size_t)img * ob.H * ob.Xp + kx;
    const C2 *K = ob.khat + ((size_t)(ob.khat_shared ? img - s * ob.C : img) * ob.Fy + ky0) * ob.Xp + kx;
    C2 *Pout = ob.P + ((size_t)img * ob.Fy + ky0) * ob.Xp + kx;
    for (int ib = 0; ib < ob.H; ib += SB_RS_QROWS) {
        C2 u[SB_RS_QROWS];
#pragma unroll
        for (int r = 0; r < SB_RS_QROWS; ++r) u[r] = ib + r < ob.H ? U[(size_t)(ib + r) * ob.Xp] : C2{T(0), T(0)};
        const int nr = min(SB_RS_QROWS, ob.H - ib);
#pragma unroll 1
        for (int q = 0; q < nky; q += 4) { // (SB_RS_KY is a multiple of 4; entries beyond nky are zero-filled in ey and not stored)
            C2 k[4], acc[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                k[m] = q + m < nky ? K[(size_t)(q + m) * ob.Xp] : C2{T(0), T(0)};
                acc[m] = C2{T(0), T(0)};
            }
#pragma unroll
            for (int r = 0; r < SB_RS_QROWS; ++r) {
                if (r < nr) { // CTA-uniform
                    T e0[4], e1[4];
                    const T *e = reinterpret_cast<const T *>(ey + (ib + r) * SB_RS_KY + q);
                    RotVec<T>::load4(e, e0);
                    RotVec<T>::load4(e + 4, e1);
                    acc[0].x += e0[0] * u[r].x - e0[1] * u[r].y, acc[0].y += e0[0] * u[r].y + e0[1] * u[r].x;
                    acc[1].x += e0[2] * u[r].x - e0[3] * u[r].y, acc[1].y += e0[2] * u[r].y + e0[3] * u[r].x;
                    acc[2].x += e1[0] * u[r].x - e1[1] * u[r].y, acc[2].y += e1[0] * u[r].y + e1[1] * u[r].x;
                    acc[3].x += e1[2] * u[r].x - e1[3] * u[r].y, acc[3].y += e1[2] * u[r].y + e1[3] * u[r].x;
                }
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (q + m < nky) {
                    C2 o;
                    o.x = ob.h2 * (k[m].x * acc[m].x - k[m].y * acc[m].y), o.y = ob.h2 * (k[m].x * acc[m].y + k[m].y * acc[m].x);
                    C2 *d = Pout + (size_t)(q + m) * ob.Xp;
                    if (ib == 0)
                        *d = o;
                    else
                        d->x += o.x, d->y += o.y;
                }
            }
        }
    }
}

// ======================================================================================================
// Rotated resampling observation (renderer.py:318-363, 498-524).  The reference shifts the kernel along both axes to every
// low-resolution row and the model along both axes to every column, then contracts the two tables.  Each two-axis shift is
// a Fourier multiplier (interpolation.shift_multiplier: the real inverse transform keeps the Hermitian part of the phase
// ramp, which matters on the Nyquist lines), so with P = K^ conj(M^)
//     LR[c,i,j] = h^2 sum_kx w_kx Re sum_ky P[c,ky,kx] A_i[ky,kx] B_j[ky,kx]                 (w = 1, 2, ..., 2, 1)
// -- the value of a band-limited image at the rotated position of pixel (i, j) -- and the adjoint is
//     G = IDFT2( h^2 K^ sum_ij R[i,j] A_i B_j ).
// Both are dense contractions over the flattened half plane k = (ky, kx): 2 H W Fy (Fx/2+1) complex multiply-adds per band
// (what the reference spends in its (H x Fy Fx)(Fy Fx x W) matrix product).  H, W <= 32.
// ======================================================================================================
#define SB_ROT_SUB 32    // k entries staged per pass of the forward contraction
#define SB_ROT_LD 40     // row pitch of the staged planes: 32 outputs + 8, so that (4 k) x (8 rows) stores hit 32 banks
#define SB_ROT_MAXB 5    // bands per CTA of the forward contraction (one warp each)
template <typename T> __device__ __forceinline__ void cp_async_c2(typename Cx<T>::type *dst, const typename Cx<T>::type *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// grid (n_chunk, ceil(C / SB_ROT_MAXB), S), 32 SB_ROT_MAXB threads.  Warp b owns band c0 + b and the whole H x W output of it over this CTA's k
// range: lane (ti, tj) = (lane / 8, lane % 8) accumulates the 8 x 4 block of rows 8 ti.. and columns 4 tj.. in registers.  The
// operands are staged per 32 k as planar real / imaginary [k][32] tiles -- the column table B_j once for all bands, and
// U = w_kx P A_i per band -- so one k costs a lane six 128-bit shared loads for 64 fused multiply-adds.
template <typename T> __global__ void __launch_bounds__(32 * SB_ROT_MAXB) k_rot_partial(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    extern __shared__ __align__(16) unsigned char smem[];
    const SpecObs<T> &ob = a.ob;
    constexpr int cb = SB_ROT_MAXB;
    const int chunk = blockIdx.x, s = blockIdx.z, c0 = blockIdx.y * cb, nb = min(cb, ob.C - c0);
    if (a.done[s]) return;
    const int H = ob.H, W = ob.W, Xp = ob.Xp, Fxc = ob.Fxc, K = ob.Fy * Xp, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int PLANE = SB_ROT_SUB * SB_ROT_LD;
    T *Br = reinterpret_cast<T *>(smem), *Bi = Br + PLANE, *Ur = Bi + PLANE, *Ui = Ur + (size_t)cb * PLANE;
    const int ti = lane >> 3, tj = lane & 7;
    T acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = T(0);
    const C2 *P = ob.P + (size_t)(s * ob.C + c0) * K;
    C2 *Praw = reinterpret_cast<C2 *>(Ui + (size_t)cb * PLANE); // [2][cb][SB_ROT_SUB] spectra of the next / current pass
    const bool even = (ob.Fx & 1) == 0;
    const int kbeg = chunk * ob.chunk, kend = min(K, kbeg + ob.chunk);
    // A thread stages elements e = tid + n blockDim of the (32 k) x (32 rows) tile, e = [q_hi:3][row_hi:2][row_lo:3][q_lo:2]: a
    // warp covers 8 table rows x 4 consecutive k (eight full 32-byte sectors) and its stores hit 32 distinct banks.  The
    // table entries of the NEXT pass are fetched into registers (and its spectra by cp.async) before the products of the
    // current one, so the loads fly under the arithmetic.
    constexpr int NT = 32 * SB_ROT_MAXB, NE = (SB_ROT_SUB * 32 + NT - 1) / NT; // elements per thread
    C2 av[NE], bv[NE];
    auto fetch = [&](int k0, int buf) {
#pragma unroll
        for (int n = 0; n < NE; ++n) {
            const int e = tid + n * NT;
            av[n] = bv[n] = C2{T(0), T(0)};
            if (e < SB_ROT_SUB * 32) {
                const int q = (e & 3) | ((e >> 7) << 2), row = ((e >> 2) & 7) | (((e >> 5) & 3) << 3), k = k0 + q;
                if (k < kend && row < W) bv[n] = ob.RB[(size_t)row * K + k];
                if (k < kend && row < H) av[n] = ob.RA[(size_t)row * K + k];
            }
        }
        for (int e = tid; e < nb * SB_ROT_SUB; e += NT) {
            const int b = e / SB_ROT_SUB, q = e - b * SB_ROT_SUB;
            if (k0 + q < kend) cp_async_c2<T>(Praw + (buf * cb + b) * SB_ROT_SUB + q, P + (size_t)b * K + k0 + q);
            else Praw[(buf * cb + b) * SB_ROT_SUB + q] = C2{T(0), T(0)};
        }
        cp_async_commit();
    };
    fetch(kbeg, 0);
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += SB_ROT_SUB, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads(); // spectra of this pass landed; every warp is done with the planes of the previous pass
#pragma unroll
        for (int n = 0; n < NE; ++n) {
            const int e = tid + n * NT;
            if (e < SB_ROT_SUB * 32) {
                const int q = (e & 3) | ((e >> 7) << 2), row = ((e >> 2) & 7) | (((e >> 5) & 3) << 3), kx = (k0 + q) % Xp;
                const T w = kx >= Fxc ? T(0) : ((kx == 0 || (even && kx == Fxc - 1)) ? T(1) : T(2));
                Br[q * SB_ROT_LD + row] = bv[n].x, Bi[q * SB_ROT_LD + row] = bv[n].y;
                const C2 m = C2{w * av[n].x, w * av[n].y};
                for (int b = 0; b < nb; ++b) {
                    const C2 p = Praw[(buf * cb + b) * SB_ROT_SUB + q];
                    Ur[b * PLANE + q * SB_ROT_LD + row] = p.x * m.x - p.y * m.y;
                    Ui[b * PLANE + q * SB_ROT_LD + row] = p.x * m.y + p.y * m.x;
                }
            }
        }
        __syncthreads();
        if (k0 + SB_ROT_SUB < kend) fetch(k0 + SB_ROT_SUB, buf ^ 1);
        if (warp < nb) {
            const T *ur = Ur + warp * PLANE + 8 * ti, *ui = Ui + warp * PLANE + 8 * ti, *br = Br + 4 * tj, *bi = Bi + 4 * tj;
#pragma unroll 4
            for (int q = 0; q < SB_ROT_SUB; ++q) {
                T u0[4], u1[4], v0[4], v1[4], x[4], y[4];
                RotVec<T>::load4(ur + q * SB_ROT_LD, u0);
                RotVec<T>::load4(ur + q * SB_ROT_LD + 4, u1);
                RotVec<T>::load4(ui + q * SB_ROT_LD, v0);
                RotVec<T>::load4(ui + q * SB_ROT_LD + 4, v1);
                RotVec<T>::load4(br + q * SB_ROT_LD, x);
                RotVec<T>::load4(bi + q * SB_ROT_LD, y);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        acc[r][m] += u0[r] * x[m];
                        acc[r][m] -= v0[r] * y[m];
                        acc[r + 4][m] += u1[r] * x[m];
                        acc[r + 4][m] -= v1[r] * y[m];
                    }
            }
        }
    }
    if (warp < nb) {
        T *out = ob.Rpart + ((size_t)(s * ob.C + c0 + warp) * ob.n_chunk + chunk) * H * W;
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int i = 8 * ti + r, j = 4 * tj + m;
                if (i < H && j < W) out[i * W + j] = acc[r][m];
            }
    }
}

// one CTA per (scene, band): chunks summed in double (fixed order) -> LR, residual r = w (LR - d), chi^2 partial
template <typename T> __global__ void __launch_bounds__(256) k_rot_residual(const SpecArgs<T> a) {
    __shared__ double red[40];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.x, s = img / ob.C;
    if (a.done[s]) return;
    const int HW = ob.H * ob.W;
    const T *part = ob.Rpart + (size_t)img * ob.n_chunk * HW;
    double chi = 0.0;
    for (int idx = threadIdx.x; idx < HW; idx += blockDim.x) {
        double acc = 0.0;
        for (int q = 0; q < ob.n_chunk; ++q) acc += (double)part[(size_t)q * HW + idx];
        const T m = (T)((double)ob.h2 * acc);
        const size_t di = (size_t)img * HW + idx;
        const T w = ob.weights[di], diff = m - ob.data[di];
        ob.Rres[di] = w * diff;
        chi += (double)w * (double)diff * (double)diff;
        if (a.rendered_out) a.rendered_out[di] = m;
    }
    chi = block_sum(chi, red);
    if (threadIdx.x == 0) a.partials[img] = chi;
}

// adjoint: P[k] = h^2 K^[k] sum_i A_i[k] sum_j R[i,j] B_j[k]; grid (ceil(Fy Xp / 128), S C), one k per thread
template <typename T> __global__ void __launch_bounds__(128) k_rot_adjoint(const SpecArgs<T> a) {
    typedef typename Cx<T>::type C2;
    extern __shared__ __align__(16) unsigned char smem[];
    const SpecObs<T> &ob = a.ob;
    const int img = blockIdx.y, s = img / ob.C;
    if (a.done[s]) return;
    const int H = ob.H, W = ob.W, K = ob.Fy * ob.Xp;
    T *R = reinterpret_cast<T *>(smem); // [H][32], columns beyond W zero: rows are read four entries per shared load
    for (int idx = threadIdx.x; idx < H * 32; idx += blockDim.x) {
        const int i = idx >> 5, j = idx & 31;
        R[idx] = j < W ? ob.Rres[(size_t)img * H * W + i * W + j] : T(0);
    }
    __syncthreads();
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= K) return;
    C2 b[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = j < W ? ob.RB[(size_t)j * K + k] : C2{T(0), T(0)};
    C2 acc = C2{T(0), T(0)};
    for (int i = 0; i < H; ++i) {
        C2 v = C2{T(0), T(0)};
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            T r[4];
            RotVec<T>::load4(R + i * 32 + j, r);
#pragma unroll
            for (int m = 0; m < 4; ++m) v.x += r[m] * b[j + m].x, v.y += r[m] * b[j + m].y;
        }
        const C2 m = ob.RA[(size_t)i * K + k];
        acc.x += m.x * v.x - m.y * v.y;
        acc.y += m.x * v.y + m.y * v.x;
    }
    const C2 kh = ob.khat[(size_t)(ob.khat_shared ? img - s * ob.C : img) * K + k];
    C2 out;
    out.x = ob.h2 * (kh.x * acc.x - kh.y * acc.y);
    out.y = ob.h2 * (kh.x * acc.y + kh.y * acc.x);
    ob.P[(size_t)img * K + k] = out;
}

// ---- dispatch table ----------------------------------------------------------------------------------
// supported transform lengths L = R1 * R2 (both the row length Fx and the column length Fy must be in this list)
#define SB_SPEC_LENGTHS(X) \
    X(6, 8) X(8, 8) X(8, 9) X(8, 10) X(8, 12) X(9, 12) X(8, 16) X(12, 12) X(10, 16) X(12, 16) X(15, 16) X(16, 16) X(16, 18) X(15, 20) X(16, 20) X(16, 24)

template <typename T> struct SpecKernels {
    typedef void (*fn)(const SpecArgs<T>);
    int R1 = 0, R2 = 0, NBcol = 0;
    typedef void (*fn_tma)(const SpecArgs<T>, const CUtensorMap, const CUtensorMap);
    fn render2 = nullptr; // render with two pixels per source and thread held ahead
    fn render = nullptr, residual = nullptr, residual_r = nullptr, grad = nullptr, column = nullptr, column_fwd = nullptr, column_inv = nullptr;
    fn_tma column_tma = nullptr; // float only
    size_t sf = 0; // Plan2::SF
};
template <typename T> struct SpecColNB { static const int value = sizeof(T) == 4 ? 16 : 8; };

// defined in spectral_f32.cu / spectral_f64.cu
bool spec_kernels_f32(int L, SpecKernels<float> *out);
bool spec_kernels_f64(int L, SpecKernels<double> *out);
int spec_supported_length(int need); // smallest supported L >= need, 0 if none

template <typename T> inline bool spec_kernels(int L, SpecKernels<T> *out);
template <> inline bool spec_kernels<float>(int L, SpecKernels<float> *out) { return spec_kernels_f32(L, out); }
template <> inline bool spec_kernels<double>(int L, SpecKernels<double> *out) { return spec_kernels_f64(L, out); }

} // namespace sb
