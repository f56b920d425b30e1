// float64 instantiations of the fused spectral kernels (separate translation unit: compiled in parallel).
#include "spectral.cuh"

namespace sb {

bool spec_kernels_f64(int L, SpecKernels<double> *out) {
#define X(A, B)                                                                   \
    if (L == (A) * (B)) {                                                         \
        out->R1 = A, out->R2 = B, out->NBcol = SpecColNB<double>::value;           \
        out->render = k_spec_render<double, A, B>;                                 \
        out->render2 = k_spec_render<double, A, B, 2>; \
        out->residual = k_spec_residual<double, A, B>;                                 \
        out->residual_r = k_spec_residual<double, A, B, true>;                             \
        out->grad = k_spec_grad<double, A, B>;                                     \
        out->column = k_spec_column<double, A, B, SpecColNB<double>::value>;        \
        out->column_fwd = k_spec_column_fwd<double, A, B, SpecColNB<double>::value>; \
        out->column_inv = k_spec_column_inv<double, A, B, SpecColNB<double>::value>; \
        out->sf = sbfft::Plan2<A, B>::SF;                                         \
        return true;                                                              \
    }
    SB_SPEC_LENGTHS(X)
#undef X
    return false;
}

} // namespace sb
