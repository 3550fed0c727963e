// Target-space charts: intrinsic <-> embedding coordinates of every sub-pdf of a pdf in one pass (one thread per row).
// Reference: main/default.py:1737-1813 (`pdf.transform_target_space`), per sub-pdf
// layers/spheres/sphere_base.py:242-332 and :796-841; Euclidean and interval sub-pdfs are identities
// (euclidean_base.py:116-118, interval_base.py:104-106).  This is what `force_embedding_coordinates` applies before
// the log_pdf chain / after the sampling chain (main/default.py:906-913, :1522-1529).
#pragma once
#include "s2.cuh"

namespace jf {

template <typename T>
struct ChartArgs {
    int n_sub, to_embedding;
    int kind[JF_MAX_SUBPDFS];      // 0: identity (copy dim columns), 1: S1, 2: S2
    int dim[JF_MAX_SUBPDFS];
    int in_col[JF_MAX_SUBPDFS], out_col[JF_MAX_SUBPDFS];
    const T* in;  int64_t ld_in;
    T* out;       int64_t ld_out;
    const T* logdet_in; T* logdet_out;
    int64_t B;
};

template <typename T>
__global__ void __launch_bounds__(256) chart_kernel(const __grid_constant__ ChartArgs<T> a) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.B) return;
    const T* x = a.in + row * a.ld_in;
    T* y = a.out + row * a.ld_out;
    T logdet = a.logdet_in ? a.logdet_in[row] : T(0);
    for (int k = 0; k < a.n_sub; ++k) {
        const T* xi = x + a.in_col[k];
        T* yo = y + a.out_col[k];
        if (a.kind[k] == 0) {
            for (int j = 0; j < a.dim[k]; ++j) yo[j] = xi[j];
        } else if (a.kind[k] == 1) {
            if (a.to_embedding) {                                  // sphere_base.py:306-311
                T s, c;
                sincos(xi[0], &s, &c);
                yo[0] = c; yo[1] = s;
            } else {                                               // sphere_base.py:259-266
                T ang = acos(xi[0] / sqrt(xi[0] * xi[0] + xi[1] * xi[1]));
                if (xi[1] < T(0)) ang = T(2 * kPi) - ang;
                yo[0] = ang;
            }
        } else {
            if (a.to_embedding) {
                T e[3];
                s2_to_embedding(xi[0], xi[1], e, logdet);
                yo[0] = e[0]; yo[1] = e[1]; yo[2] = e[2];
            } else {
                const T e[3] = {xi[0], xi[1], xi[2]};
                T theta, phi;
                s2_from_embedding(e, theta, phi, logdet);
                yo[0] = theta; yo[1] = phi;
            }
        }
    }
    if (a.logdet_out) a.logdet_out[row] = logdet;
}

// out[r] = log( mean_c exp(in[r, c]) ): the S x S cross-evaluation reduce of the marginal entropies
// (reference main/default.py:2444-2448).  One warp per row, two passes (max, then rescaled sum), coalesced reads.
template <typename T>
__global__ void __launch_bounds__(256) row_logmeanexp_kernel(const T* in, int64_t rows, int64_t cols, T* out) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const T* p = in + row * cols;
    T m = -Num<T>::big;
    for (int64_t c = lane; c < cols; c += 32) m = tmax(m, p[c]);
    for (int o = 16; o > 0; o >>= 1) m = tmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    T s = 0;
    for (int64_t c = lane; c < cols; c += 32) s += exp(p[c] - m);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[row] = m + log(s) - log(T(cols));
}

}  // namespace jf
