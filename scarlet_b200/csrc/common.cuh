// Shared device/host definitions of the scarlet_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "scarlet_b200.h"

#define SB_MAXC SB_MAX_CHANNELS

namespace sb {

// ---- error plumbing -------------------------------------------------------------------------------
extern thread_local std::string g_err;
int set_err(int code, const char *fmt, ...);

#define SB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return sb::set_err(SB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define SB_CUFFT(call)                                                                             \
    do {                                                                                           \
        cufftResult r_ = (call);                                                                   \
        if (r_ != CUFFT_SUCCESS)                                                                   \
            return sb::set_err(SB_ERR_CUFFT, "%s:%d %s -> cufft error %d", __FILE__, __LINE__, #call, (int)r_); \
    } while (0)
#define SB_TRY(call)           \
    do {                       \
        int rc_ = (call);      \
        if (rc_ != SB_OK) return rc_; \
    } while (0)

// ---- real/complex type traits ---------------------------------------------------------------------
template <typename T> struct Cx;
template <> struct Cx<float> {
    typedef float2 type;
    typedef cufftReal real_t;
    typedef cufftComplex cplx_t;
    static const cufftType r2c = CUFFT_R2C, c2r = CUFFT_C2R;
    static cufftResult fwd(cufftHandle p, float *in, float2 *out) { return cufftExecR2C(p, in, out); }
    static cufftResult inv(cufftHandle p, float2 *in, float *out) { return cufftExecC2R(p, in, out); }
};
template <> struct Cx<double> {
    typedef double2 type;
    typedef cufftDoubleReal real_t;
    typedef cufftDoubleComplex cplx_t;
    static const cufftType r2c = CUFFT_D2Z, c2r = CUFFT_Z2D;
    static cufftResult fwd(cufftHandle p, double *in, double2 *out) { return cufftExecD2Z(p, in, out); }
    static cufftResult inv(cufftHandle p, double2 *in, double *out) { return cufftExecZ2D(p, in, out); }
};

// ---- device-side descriptors ----------------------------------------------------------------------
struct DevSource {
    int kind, By, Bx, oy, ox, chain, sed_chain, sed_is_f32, morph_fixed, sed_fixed;
    int scene, point_idx; // point_idx: slot in the centre arrays (point-source centre, or sub-pixel shift of a shifting image)
    int shifting, shift_Fy, shift_Fx, resizing; // Fourier-shifted image morphology (fft.py:399-428) on its own fast grid; dynamic box
    long long toep_off;  // offset (in vectors of length 2*Bmax-1) of this source's 8 Toeplitz vectors
    double shift_step;
    long long morph_off; // kind 0: element offset in the packed morphology arrays; kind 1: offset in pmorph
    double morph_step, sed_step_factor;
    double sed_step_min[SB_MAXC];
};

// Wavefront image of one radial-monotonicity operator.  Tasks (= entries of dist_idx) are sorted by
// dependency level; a task's positive-weight neighbours are kept in the reference's offset order.
struct DevMono {
    int n_pix, n_tasks, n_levels, nb; // nb = neighbour slots per task (4 or 8)
    int off[8];
    const int *pix;           // [n_tasks]
    const unsigned *code;     // [n_tasks] 4 bits per slot: offset index, 15 = empty
    const void *w;            // T [nb][n_tasks]
    const int *level_start;   // [n_levels+1]
    // the same operator laid out by TRIPS for the warp-per-source kernel (update_warp.cuh), nb == 4 only: every level cut
    // into trips of exactly 32 slots, dummy tasks in the unused slots, one dummy trip behind the last; one byte image
    // [W4<T> x w_cap | uint2 x w_cap | u16 x w_cap] that the kernel pulls into shared memory with bulk async copies
    const void *wtab;
    int w_trips, w_cap;       // trips (even), entries = 32 (w_trips + 1)
};

struct DevChain {
    int n_ops, repeat;
    sb_op ops[SB_MAX_CHAIN_OPS];
};

template <typename T> struct DevObs {
    int kind, C, H, W, chan_off, oy, ox, Fy, Fx, Fxc, khat_shared;
    int Kp;                        // row pitch of khat (complex elements)
    int Bh, Bw;                    // rows per image and row pitch of B: (Fy, Fx) on the cuFFT path, (Ny, Nx) on the fused path
    T *A;                          // [S][C][Fy][Fx] real grid: model in, residual in (pad region stays zero)
    T *B;                          // [S][C][Bh][Bw] convolution out: rendered model, later the gradient wrt the model
    typename Cx<T>::type *Ahat;    // [S][C][Fy][Fxc]
    typename Cx<T>::type *khat;    // [S or 1][C][Fy][Fxc], 1/(Fy*Fx) folded in
    const T *data, *weights;       // [S][C][H][W]
};

struct FitScalars {
    int prox_max_iter, min_iter, fixed_iterations, overwrite_vhat_at_it0, pause_every, _pad;
    double e_rel, b1, b2, eps;
};

// ---- small device helpers -------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum / max, result broadcast to every thread.  sh: >= 33 doubles.  Deterministic for a fixed
// block size (fixed tree), which keeps repeated runs bit-identical.
__device__ __forceinline__ double block_sum(double v, double *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? sh[lane] : 0.0;
        v = warp_sum(v);
        if (lane == 0) sh[32] = v;
    }
    __syncthreads();
    v = sh[32];
    __syncthreads();
    return v;
}
__device__ __forceinline__ double block_max(double v, double *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? sh[lane] : -INFINITY;
        v = warp_max(v);
        if (lane == 0) sh[32] = v;
    }
    __syncthreads();
    v = sh[32];
    __syncthreads();
    return v;
}

// un-fused arithmetic: the monotonic sweep must round like the reference's scalar C++ loop
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

} // namespace sb
