// CPU harness for scarlet_b200/csrc/fft_core.cuh: emulates the lanes of the two-stage transform with plain loops so that
// the in-register butterflies, twiddles and the shared-memory exchange can be checked against numpy.fft without a GPU.
// TEST INFRASTRUCTURE ONLY (built by tests/test_fft_core.py with g++).
#include <vector>

#include "fft_core.cuh"

using namespace sbfft;

template <typename T, int R1, int R2> static void run2(const T *in, T *out, int inverse) {
    typedef typename CpxOf<T>::type C;
    typedef Plan2<R1, R2> P;
    constexpr int L = P::L;
    std::vector<C> tw(L), sm(P::SF);
    for (int k1 = 0; k1 < R1; ++k1)
        for (int n2 = 0; n2 < R2; ++n2) {
            const TwPair t = ct_twiddle(n2 * k1, L);
            tw[k1 * R2 + n2] = C{(T)t.c, (T)(-t.s)};
        }
    if (!inverse) {
        for (int n2 = 0; n2 < R2; ++n2) { // stage A lanes
            C a[R1];
            for (int n1 = 0; n1 < R1; ++n1) a[n1] = C{in[2 * (n1 * R2 + n2)], in[2 * (n1 * R2 + n2) + 1]};
            fwd_stage_a<R1, R2>(a, n2, tw.data(), sm.data());
        }
        for (int k1 = 0; k1 < R1; ++k1) { // stage B lanes
            C b[R2];
            fwd_stage_b<R1, R2>(b, k1, sm.data());
            for (int k2 = 0; k2 < R2; ++k2) out[2 * (k1 + R1 * k2)] = b[k2].x, out[2 * (k1 + R1 * k2) + 1] = b[k2].y;
        }
    } else {
        for (int k1 = 0; k1 < R1; ++k1) {
            C b[R2];
            for (int k2 = 0; k2 < R2; ++k2) b[k2] = C{in[2 * (k1 + R1 * k2)], in[2 * (k1 + R1 * k2) + 1]};
            inv_stage_b<R1, R2>(b, k1, tw.data(), sm.data());
        }
        for (int n2 = 0; n2 < R2; ++n2) {
            C a[R1];
            inv_stage_a<R1, R2>(a, n2, sm.data());
            for (int n1 = 0; n1 < R1; ++n1) out[2 * (n1 * R2 + n2)] = a[n1].x, out[2 * (n1 * R2 + n2) + 1] = a[n1].y;
        }
    }
}

template <typename T, int N> static void runreg(const T *in, T *out, int inverse) {
    typedef typename CpxOf<T>::type C;
    C x[N];
    for (int i = 0; i < N; ++i) x[i] = C{in[2 * i], in[2 * i + 1]};
    if (inverse)
        regfft<N, true>(x);
    else
        regfft<N, false>(x);
    for (int i = 0; i < N; ++i) out[2 * i] = x[i].x, out[2 * i + 1] = x[i].y;
}

#define SB_FFT_LENGTHS(X) X(6, 8) X(8, 8) X(8, 9) X(8, 10) X(8, 12) X(10, 10) X(10, 12) X(8, 16) X(12, 12) X(10, 15) X(10, 16) \
    X(12, 15) X(12, 16) X(10, 20) X(12, 18) X(15, 16) X(16, 16) X(16, 18) X(15, 20) X(16, 20) X(18, 18) X(18, 20) X(16, 24) X(20, 20)

template <typename T> static int dispatch2(int R1, int R2, const T *in, T *out, int inverse) {
#define X(A, B)                          \
    if (R1 == A && R2 == B) {            \
        run2<T, A, B>(in, out, inverse); \
        return 0;                        \
    }
    SB_FFT_LENGTHS(X)
#undef X
    return -1;
}
template <typename T> static int dispatchreg(int N, const T *in, T *out, int inverse) {
#define R(A)                            \
    if (N == A) {                       \
        runreg<T, A>(in, out, inverse); \
        return 0;                       \
    }
    R(2) R(3) R(4) R(5) R(6) R(8) R(9) R(10) R(12) R(15) R(16) R(18) R(20) R(24) R(25)
#undef R
    return -1;
}

extern "C" {
int fftcore_two_stage_f32(int R1, int R2, const float *in, float *out, int inverse) { return dispatch2<float>(R1, R2, in, out, inverse); }
int fftcore_two_stage_f64(int R1, int R2, const double *in, double *out, int inverse) { return dispatch2<double>(R1, R2, in, out, inverse); }
int fftcore_reg_f32(int N, const float *in, float *out, int inverse) { return dispatchreg<float>(N, in, out, inverse); }
int fftcore_reg_f64(int N, const double *in, double *out, int inverse) { return dispatchreg<double>(N, in, out, inverse); }
}
