/* CPU oracle: sequential radial-monotonicity sweep.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the algorithm of the reference's native operator
 * (scarlet/operators_pybind11.cc:14-36, bound for float and double at :243-246):
 * walk the pixels in order of increasing distance from the centre (dist_idx, centre
 * excluded by the caller, scarlet/operator.py:92) and clamp each pixel to
 * (1 - min_gradient) times the weighted sum of its strictly-closer neighbours.
 * weights is row-major (n_off, n_pix); only weights > 0 contribute, in offset order.
 */
#include <stddef.h>

#define SWEEP(NAME, T)                                                                      \
    void NAME(T *img, const T *weights, const int *offsets, int n_off, const int *dist_idx, \
              int n_idx, int n_pix, T min_gradient)                                         \
    {                                                                                       \
        const T keep = (T)1 - min_gradient;                                                 \
        for (int d = 0; d < n_idx; ++d) {                                                   \
            const int p = dist_idx[d];                                                      \
            T ref = 0;                                                                      \
            for (int i = 0; i < n_off; ++i) {                                               \
                const T w = weights[(size_t)i * n_pix + p];                                 \
                if (w > 0) ref += img[p + offsets[i]] * w;                                  \
            }                                                                               \
            const T cap = ref * keep;                                                       \
            if (cap < img[p]) img[p] = cap;                                                 \
        }                                                                                   \
    }

SWEEP(oracle_monotonic_sweep_f64, double)
SWEEP(oracle_monotonic_sweep_f32, float)
