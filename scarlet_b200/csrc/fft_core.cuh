// In-register mixed-radix FFT building blocks for the fused spectral-convolution kernels (spectral.cuh).
//
// A length-L transform is split as L = R1 * R2 ("four-step"): every participating thread runs an R1-point and
// later an R2-point FFT entirely in registers (regfft<N>, recursive Cooley-Tukey over the radices 2/3/4/5, all
// indices and twiddle factors compile-time constants); the two halves exchange data once through shared memory.
// The functions are __host__ __device__ so that the arithmetic can be unit-tested on the CPU (tests/fftcore/).
//
// This replaces, for the per-iteration convolutions, what the reference does with numpy.fft.rfftn / irfftn in
// scarlet/fft.py:255-273 (Fourier.fft) and 200-243 (Fourier.from_fft).
#pragma once
#include <type_traits>
#include <utility>

#ifdef __CUDACC__
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD inline
struct float2 {
    float x, y;
};
struct double2 {
    double x, y;
};
#endif

namespace sbfft {

template <typename T> struct CpxOf;
template <> struct CpxOf<float> { typedef float2 type; };
template <> struct CpxOf<double> { typedef double2 type; };

// ---- compile-time loop -------------------------------------------------------------------------------
template <int I, int End, typename F> SB_HD void static_for(F &&f) {
    if constexpr (I < End) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, End>(static_cast<F &&>(f));
    }
}

// ---- compile-time trigonometry: cos / sin of 2*pi*j/N, exact octant reduction, Taylor series on [0, pi/4] ---------
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double ct_sin(double x) {
    double x2 = x * x, term = x, sum = x;
    for (int n = 1; n <= 12; ++n) {
        term *= -x2 / ((2.0 * n) * (2.0 * n + 1.0));
        sum += term;
    }
    return sum;
}
constexpr double ct_cos(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int n = 1; n <= 12; ++n) {
        term *= -x2 / ((2.0 * n - 1.0) * (2.0 * n));
        sum += term;
    }
    return sum;
}
struct TwPair {
    double c, s;
};
constexpr TwPair ct_twiddle(int j, int N) { // (cos, sin) of 2*pi*j/N
    j %= N;
    if (j < 0) j += N;
    const int q = (8 * j) / N, r = 8 * j - q * N;
    const double a = (kPi / 4) * r / N, b = (kPi / 4) * (N - r) / N;
    const double Ca = ct_cos(a), Sa = ct_sin(a), Cb = ct_cos(b), Sb = ct_sin(b);
    switch (q) {
    case 0: return {Ca, Sa};
    case 1: return {Sb, Cb};
    case 2: return {-Sa, Ca};
    case 3: return {-Cb, Sb};
    case 4: return {-Ca, -Sa};
    case 5: return {-Sb, -Cb};
    case 6: return {Sa, -Ca};
    default: return {Cb, -Sb};
    }
}
template <int J, int N> struct Tw {
    static constexpr double c = ct_twiddle(J, N).c;
    static constexpr double s = ct_twiddle(J, N).s;
};

// ---- complex helpers -----------------------------------------------------------------------------------
template <typename C> SB_HD C cadd(C a, C b) { return C{a.x + b.x, a.y + b.y}; }
template <typename C> SB_HD C csub(C a, C b) { return C{a.x - b.x, a.y - b.y}; }
template <typename C> SB_HD C cmul(C a, C b) { return C{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename C> SB_HD C cmul_conj(C a, C b) { return C{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; } // a * conj(b)
// multiply by -i (forward) or +i (inverse)
template <bool INV, typename C> SB_HD C mul_mi(C a) {
    if constexpr (INV)
        return C{-a.y, a.x};
    else
        return C{a.y, -a.x};
}
// a * exp(-+ 2 pi i J / N) with a compile-time twiddle (sign - forward, + inverse)
template <int J, int N, bool INV, typename C> SB_HD C mul_tw(C a) {
    typedef decltype(a.x) T;
    constexpr int j = ((J % N) + N) % N;
    if constexpr (j == 0)
        return a;
    else if constexpr (2 * j == N)
        return C{-a.x, -a.y};
    else if constexpr (4 * j == N)
        return mul_mi<INV>(a);
    else if constexpr (4 * j == 3 * N) {
        C t = mul_mi<INV>(a);
        return C{-t.x, -t.y};
    } else {
        constexpr T c = (T)Tw<j, N>::c;
        constexpr T s = INV ? (T)Tw<j, N>::s : (T)(-Tw<j, N>::s);
        return C{a.x * c - a.y * s, a.x * s + a.y * c};
    }
}

// ---- base butterflies -------------------------------------------------------------------------------
template <bool INV, typename C> SB_HD void dft2(C (&x)[2]) {
    const C a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
}
template <bool INV, typename C> SB_HD void dft3(C (&x)[3]) {
    typedef decltype(x[0].x) T;
    constexpr T s = (T)0.86602540378443864676372317075294; // sin(2 pi / 3)
    const C t1 = cadd(x[1], x[2]);
    const C m = C{x[0].x - (T)0.5 * t1.x, x[0].y - (T)0.5 * t1.y};
    const C d0 = csub(x[1], x[2]);
    const C d = mul_mi<INV>(C{s * d0.x, s * d0.y});
    x[0] = cadd(x[0], t1);
    x[1] = cadd(m, d);
    x[2] = csub(m, d);
}
template <bool INV, typename C> SB_HD void dft4(C (&x)[4]) {
    const C t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]), t2 = cadd(x[1], x[3]);
    const C t3 = mul_mi<INV>(csub(x[1], x[3]));
    x[0] = cadd(t0, t2);
    x[2] = csub(t0, t2);
    x[1] = cadd(t1, t3);
    x[3] = csub(t1, t3);
}
template <bool INV, typename C> SB_HD void dft5(C (&x)[5]) {
    typedef decltype(x[0].x) T;
    constexpr T c1 = (T)Tw<1, 5>::c, c2 = (T)Tw<2, 5>::c, s1 = (T)Tw<1, 5>::s, s2 = (T)Tw<2, 5>::s;
    const C t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
    const C a1 = C{x[0].x + c1 * t1.x + c2 * t2.x, x[0].y + c1 * t1.y + c2 * t2.y};
    const C a2 = C{x[0].x + c2 * t1.x + c1 * t2.x, x[0].y + c2 * t1.y + c1 * t2.y};
    const C b1 = mul_mi<INV>(C{s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y});
    const C b2 = mul_mi<INV>(C{s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y});
    x[0] = cadd(x[0], cadd(t1, t2));
    x[1] = cadd(a1, b1);
    x[4] = csub(a1, b1);
    x[2] = cadd(a2, b2);
    x[3] = csub(a2, b2);
}

template <int N> struct Radix { // first factor of the in-register decomposition N = r * (N / r)
    static constexpr int r = (N % 4 == 0) ? 4 : (N % 2 == 0) ? 2 : (N % 3 == 0) ? 3 : 5;
};

// N-point FFT on registers, natural order in and out.  INV selects exp(+i...) (unnormalised inverse).
template <int N, bool INV, typename C> SB_HD void regfft(C (&x)[N]) {
    if constexpr (N == 1) {
    } else if constexpr (N == 2)
        dft2<INV>(x);
    else if constexpr (N == 3)
        dft3<INV>(x);
    else if constexpr (N == 4)
        dft4<INV>(x);
    else if constexpr (N == 5)
        dft5<INV>(x);
    else {
        constexpr int N1 = Radix<N>::r, N2 = N / N1;
        static_assert(N1 * N2 == N && (N1 == 2 || N1 == 3 || N1 == 4 || N1 == 5), "unsupported FFT length");
        // step 1: N1-point DFTs over n1 for every n2, then the twiddle exp(-+2 pi i n2 k1 / N)
        static_for<0, N2>([&](auto n2c) {
            constexpr int n2 = decltype(n2c)::value;
            C a[N1];
            static_for<0, N1>([&](auto i) { a[decltype(i)::value] = x[decltype(i)::value * N2 + n2]; });
            regfft<N1, INV>(a);
            static_for<0, N1>([&](auto i) {
                constexpr int k1 = decltype(i)::value;
                x[k1 * N2 + n2] = mul_tw<n2 * k1, N, INV>(a[k1]);
            });
        });
        // step 2: N2-point FFTs over n2 for every k1; X[k1 + N1 k2] ends up at position k1 * N2 + k2
        C t[N];
        static_for<0, N1>([&](auto k1c) {
            constexpr int k1 = decltype(k1c)::value;
            C b[N2];
            static_for<0, N2>([&](auto i) { b[decltype(i)::value] = x[k1 * N2 + decltype(i)::value]; });
            regfft<N2, INV>(b);
            static_for<0, N2>([&](auto i) { t[k1 + N1 * decltype(i)::value] = b[decltype(i)::value]; });
        });
        static_for<0, N>([&](auto i) { x[decltype(i)::value] = t[decltype(i)::value]; });
    }
}

// ---- two-stage length-L transform, L = R1 * R2, cooperative through shared memory ---------------------------------
// Shared-memory image of one transform: R1 rows of PR (odd pitch >= R2) complex numbers; consecutive transforms are
// SF (odd) complex numbers apart, so both "transform index fastest" and "element index fastest" thread mappings are
// bank-conflict free for 8-byte accesses.
template <int R1, int R2> struct Plan2 {
    static constexpr int L = R1 * R2;
    static constexpr int PR = R2 | 1;
    static constexpr int SF = (R1 * PR) | 1;
    static constexpr int RMAX = R1 > R2 ? R1 : R2;
};

// FORWARD flow "A then B":  x[n1 * R2 + n2] in registers a[n1] of lane n2  -->  X[k1 + R1 * k2] in b[k2] of lane k1.
// tw: [R1][R2] table of exp(-2 pi i n2 k1 / L) (global or shared memory).
template <int R1, int R2, typename C> SB_HD void fwd_stage_a(C (&a)[R1], int n2, const C *tw, C *sm) {
    regfft<R1, false>(a);
    static_for<0, R1>([&](auto i) {
        constexpr int k1 = decltype(i)::value;
        if constexpr (k1 == 0)
            sm[n2] = a[0];
        else
            sm[k1 * Plan2<R1, R2>::PR + n2] = cmul(a[k1], tw[k1 * R2 + n2]);
    });
}
template <int R1, int R2, typename C> SB_HD void fwd_stage_b(C (&b)[R2], int k1, const C *sm) {
    static_for<0, R2>([&](auto i) { b[decltype(i)::value] = sm[k1 * Plan2<R1, R2>::PR + decltype(i)::value]; });
    regfft<R2, false>(b);
}
// INVERSE flow "B then A":  X[k1 + R1 * k2] in b[k2] of lane k1  -->  x[n1 * R2 + n2] in a[n1] of lane n2 (unnormalised).
template <int R1, int R2, typename C> SB_HD void inv_stage_b(C (&b)[R2], int k1, const C *tw, C *sm) {
    regfft<R2, true>(b);
    static_for<0, R2>([&](auto i) {
        constexpr int n2 = decltype(i)::value;
        sm[k1 * Plan2<R1, R2>::PR + n2] = cmul_conj(b[n2], tw[k1 * R2 + n2]);
    });
}
template <int R1, int R2, typename C> SB_HD void inv_stage_a(C (&a)[R1], int n2, const C *sm) {
    static_for<0, R1>([&](auto i) { a[decltype(i)::value] = sm[decltype(i)::value * Plan2<R1, R2>::PR + n2]; });
    regfft<R1, true>(a);
}

} // namespace sbfft
