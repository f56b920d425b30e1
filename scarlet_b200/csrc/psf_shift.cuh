// ConvolutionRenderer(psf_shift=...): the difference kernel is moved by a fitted sub-pixel offset (sm_100a).
//
// Reference: scarlet/renderer.py:172-177 (the free ``psf_shift`` parameter, step 1e-2), 220-227 (``convolve`` shifts the
// kernel image with ``fft.shift`` on every call), 252-254; scarlet/fft.py:399-428 (``shift``: centre-pad to the fast grid of
// (kernel, kernel, padding 10), rfftn, phase ramp, irfftn, centre-crop back to P x P).
//
// Restricted to the P x P kernel box the reference's shift is EXACTLY a pair of small Toeplitz products per band,
//     K_s = Re(Cy) K Re(Tx)^T - Im(Cy) K Im(Tx)^T,
// with Cy[d] = 1/Fy sum_k exp(2 pi i (k d - m_k s0)/Fy) (m_k the signed frequency: complex inverse transform along y) and
// Tx[d] = 1/Fx sum_{k<=Fx/2} c_k exp(2 pi i k (d - s1)/Fx), c = 1,2,...,2,1 (what the real inverse transform along x does,
// Nyquist row included) -- the same identity the shifting morphologies use (kernels.cuh).  Everything here runs in double
// precision: K^ is formed in double from the kernel image in the plain plan as well.
//
// The render is linear in K_s:  rendered[y, x] = sum_uv K_s[u, v] M[y - u + P/2, x - v + P/2]  ("same" convolution), hence
//     dL/dK_s[u, v] = sum_yx r[y, x] M[y - u + P/2, x - v + P/2]        (r = w (rendered - data) in the model frame)
//     dL/ds_j       = sum_c < dL/dK_s[c], dK_s[c]/ds_j >
// with dK_s/ds_j from the derivatives of the Toeplitz vectors.
#pragma once
#include "kernels.cuh"

namespace sb {

template <typename T> struct PsfShiftArgs {
    int S, C, Py, Px, Fy, Fx;  // kernel box, fast grid of fft.shift (fft.py:116-167 with padding 10)
    int Ny, Nx, chan_off, Cm;  // model frame, first model channel of this observation, model channels
    int slot0;                 // centre-array slot of scene 0's shift (scene s: slot0 + s)
    int kernel_shared;         // 1: one kernel image for all scenes
    int fixed, mode;           // fixed: parameter not fitted; mode 1: gradients only
    double step;
    const double *kimg;        // [S or 1][C][Py][Px] unshifted difference kernel
    double *ks, *kd0, *kd1;    // [S][C][Py][Px] shifted kernel and its derivatives wrt s0, s1
    double *gk;                // [S][C][Py][Px] dL/dK_s
    const T *model;            // [S][Cm][Ny][Nx]
    const T *resid;            // [S][C][Ny][Nx] residual in the model frame
    double *center, *cen_m, *cen_v, *cen_vhat, *g_center;
    const int *it_ptr, *done;
    int *status;
    FitScalars fs;
};

// one CTA per (scene, band): Toeplitz vectors for the scene's current shift, then K_s, dK_s/ds0, dK_s/ds1
template <typename T> __global__ void __launch_bounds__(128) k_psf_kernel(const PsfShiftArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int s = blockIdx.x / a.C, c = blockIdx.x - s * a.C;
    if (a.done[s]) return;
    const int Py = a.Py, Px = a.Px, ny = 2 * Py - 1, nx = 2 * Px - 1, n = Py * Px;
    double *vec = reinterpret_cast<double *>(smem); // 8 vectors: Re Cy, Im Cy, dRe Cy, dIm Cy [ny each]; Re Tx, Im Tx, dRe, dIm [nx each]
    double *u = vec + 4 * ny + 4 * nx, *w = u + n, *w2 = w + n; // kernel image, two row-pass scratch images
    const double s0 = a.center[2 * (a.slot0 + s)], s1 = a.center[2 * (a.slot0 + s) + 1];
    for (int idx = threadIdx.x; idx < ny + nx; idx += blockDim.x) {
        const bool ydir = idx < ny;
        const int F = ydir ? a.Fy : a.Fx, off = ydir ? idx - (Py - 1) : idx - ny - (Px - 1);
        const double sh = ydir ? s0 : s1, wq = 2.0 * M_PI / F;
        double re = 0, im = 0, dre = 0, dim = 0;
        if (ydir) {
            for (int k = 0; k < F; ++k) {
                const int m = k < (F + 1) / 2 ? k : k - F; // numpy.fft.fftfreq
                double sn, cs;
                sincos(wq * ((double)k * off - (double)m * sh), &sn, &cs);
                re += cs, im += sn;
                const double f = -wq * m; // d/ds0 of the phase
                dre += -f * sn, dim += f * cs;
            }
        } else {
            for (int k = 0; k <= F / 2; ++k) {
                const double ck = (k == 0 || 2 * k == F) ? 1.0 : 2.0;
                double sn, cs;
                sincos(wq * k * ((double)off - sh), &sn, &cs);
                re += ck * cs, im += ck * sn;
                const double f = -wq * k;
                dre += -ck * f * sn, dim += ck * f * cs;
            }
        }
        double *base = ydir ? vec : vec + 4 * ny;
        const int len = ydir ? ny : nx, j = ydir ? idx : idx - ny;
        base[j] = re / F, base[len + j] = im / F, base[2 * len + j] = dre / F, base[3 * len + j] = dim / F;
    }
    const double *kin = a.kimg + ((size_t)(a.kernel_shared ? 0 : s) * a.C + c) * n;
    for (int p = threadIdx.x; p < n; p += blockDim.x) u[p] = kin[p];
    __syncthreads();
    const double *RCy = vec, *ICy = vec + ny, *dRCy = vec + 2 * ny, *dICy = vec + 3 * ny;
    const double *RTx = vec + 4 * ny, *ITx = RTx + nx, *dRTx = RTx + 2 * nx, *dITx = RTx + 3 * nx;
    double *outs[3] = {a.ks, a.kd0, a.kd1};
    // out = Ay u Bx^T - A'y u B'x^T for (Ay, Bx, A'y, B'x) = (RCy, RTx, ICy, ITx), d/ds0: (dRCy, RTx, dICy, ITx), d/ds1: (RCy, dRTx, ICy, dITx)
    const double *Ay[3] = {RCy, dRCy, RCy}, *Bx[3] = {RTx, RTx, dRTx}, *Ay2[3] = {ICy, dICy, ICy}, *Bx2[3] = {ITx, ITx, dITx};
    for (int q = 0; q < 3; ++q) {
        for (int p = threadIdx.x; p < n; p += blockDim.x) { // along x
            const int y = p / Px, x1 = p - y * Px;
            double acc = 0, acc2 = 0;
            for (int x = 0; x < Px; ++x) {
                const double v = u[y * Px + x];
                acc += v * Bx[q][x1 - x + Px - 1];
                acc2 += v * Bx2[q][x1 - x + Px - 1];
            }
            w[p] = acc, w2[p] = acc2;
        }
        __syncthreads();
        double *o = outs[q] + ((size_t)s * a.C + c) * n;
        for (int p = threadIdx.x; p < n; p += blockDim.x) { // along y
            const int y1 = p / Px, x = p - y1 * Px;
            double acc = 0;
            for (int y = 0; y < Py; ++y) acc += w[y * Px + x] * Ay[q][y1 - y + Py - 1] - w2[y * Px + x] * Ay2[q][y1 - y + Py - 1];
            o[p] = acc;
        }
        __syncthreads();
    }
}

// dL/dK_s: grid (Py, C, S); a CTA computes the Px entries of kernel row u for one band of one scene
template <typename T> __global__ void __launch_bounds__(256) k_psf_corr(const PsfShiftArgs<T> a) {
    const int u = blockIdx.x, c = blockIdx.y, s = blockIdx.z;
    if (a.done[s]) return;
    const int Ny = a.Ny, Nx = a.Nx, cu = a.Py / 2, cv = a.Px / 2;
    const T *R = a.resid + ((size_t)s * a.C + c) * Ny * Nx;
    const T *M = a.model + ((size_t)s * a.Cm + a.chan_off + c) * Ny * Nx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int v = warp; v < a.Px; v += nw) {
        double acc = 0.0;
        for (int y = 0; y < Ny; ++y) {
            const int ym = y - u + cu;
            if ((unsigned)ym >= (unsigned)Ny) continue;
            const T *r = R + (size_t)y * Nx, *m = M + (size_t)ym * Nx;
            for (int x = lane; x < Nx; x += 32) {
                const int xm = x - v + cv;
                if ((unsigned)xm < (unsigned)Nx) acc += (double)r[x] * (double)m[xm];
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) a.gk[(((size_t)s * a.C + c) * a.Py + u) * a.Px + v] = acc;
    }
}

// one CTA per scene: dL/ds_j = sum_c <dL/dK_s, dK_s/ds_j>, then the AMSGrad step of the shift (no constraint)
template <typename T> __global__ void __launch_bounds__(128) k_psf_update(const PsfShiftArgs<T> a) {
    __shared__ double red[40];
    const int s = blockIdx.x;
    if (a.done[s]) return;
    const size_t n = (size_t)a.C * a.Py * a.Px, base = (size_t)s * n;
    double g0 = 0.0, g1 = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double g = a.gk[base + i];
        g0 += g * a.kd0[base + i];
        g1 += g * a.kd1[base + i];
    }
    g0 = block_sum(g0, red);
    g1 = block_sum(g1, red);
    if (threadIdx.x == 0) {
        const int slot = a.slot0 + s;
        if (a.mode == 1) {
            if (a.g_center) a.g_center[2 * slot] = g0, a.g_center[2 * slot + 1] = g1;
            return;
        }
        if (a.fixed) return;
        const int it = a.it_ptr[s];
        const double g2[2] = {g0, g1};
        double *cen = a.center + 2 * slot, *m = a.cen_m + 2 * slot, *v = a.cen_v + 2 * slot, *vh = a.cen_vhat + 2 * slot;
        for (int i = 0; i < 2; ++i) {
            double mm = m[i], vv = v[i], vvh = vh[i];
            const double psi = amsgrad(g2[i], mm, vv, vvh, it, a.fs);
            m[i] = mm, v[i] = vv, vh[i] = vvh;
            cen[i] = cen[i] - a.step * mm / psi;
            if (!isfinite(cen[i])) atomicExch(a.status + s, SB_ERR_NONFINITE);
        }
    }
}

} // namespace sb
