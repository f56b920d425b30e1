// float32 instantiations of the fused spectral kernels (separate translation unit: compiled in parallel).
#include "spectral.cuh"

namespace sb {

bool spec_kernels_f32(int L, SpecKernels<float> *out) {
#define X(A, B)                                                                   \
    if (L == (A) * (B)) {                                                         \
        out->R1 = A, out->R2 = B, out->NBcol = SpecColNB<float>::value;           \
        out->render = k_spec_render<float, A, B>;                                 \
        out->render2 = k_spec_render<float, A, B, 2>; \
        out->residual = k_spec_residual<float, A, B>;                                 \
        out->residual_r = k_spec_residual<float, A, B, true>;                             \
        out->grad = k_spec_grad<float, A, B>;                                     \
        out->column = k_spec_column<float, A, B, SpecColNB<float>::value>;        \
        out->column_tma = k_spec_column_tma<float, A, B, SpecColNB<float>::value>; \
        out->column_fwd = k_spec_column_fwd<float, A, B, SpecColNB<float>::value>; \
        out->column_inv = k_spec_column_inv<float, A, B, SpecColNB<float>::value>; \
        out->sf = sbfft::Plan2<A, B>::SF;                                         \
        return true;                                                              \
    }
    SB_SPEC_LENGTHS(X)
#undef X
    return false;
}

int spec_supported_length(int need) {
    int best = 0;
#define X(A, B) \
    if ((A) * (B) >= need && (best == 0 || (A) * (B) < best)) best = (A) * (B);
    SB_SPEC_LENGTHS(X)
#undef X
    return best;
}

} // namespace sb
