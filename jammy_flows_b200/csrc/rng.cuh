// Base-space normals on the device: Philox4x32-10 (Salmon et al., SC'11) keyed by the seed, counter = (global row index,
// pair index), two 53-bit uniforms per block, Box-Muller.  Replaces the reference's host-side
// `numpy.random.normal(size=(samplesize, total_base_dim))` + H2D copy (main/default.py:1661-1668) for pdf.sample();
// because a row depends only on (seed, global row), every rank of a sharded job draws its slice of the SAME stream:
// the union of the per-rank sample sets does not depend on the number of GPUs.
#pragma once
#include "common.cuh"

namespace jf {

JF_DEVINL void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* r) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

// 53-bit uniform in (0,1) from two words
JF_DEVINL double u53(uint32_t hi, uint32_t lo) {
    const unsigned long long m = ((((unsigned long long)hi) << 21) ^ (((unsigned long long)lo) >> 11)) & ((1ull << 53) - 1);
    return ((double)m + 0.5) * 1.1102230246251565e-16;   // 2^-53
}

// one thread per (row, pair): out[row, 2j], out[row, 2j+1]
template <typename T>
__global__ void __launch_bounds__(256) normal_rows_kernel(unsigned long long seed, unsigned long long first_row, int64_t B,
                                                          int dim, T* out, int64_t ld) {
    const int n_pairs = (dim + 1) >> 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n_pairs) return;
    const int64_t i = idx / n_pairs;
    const int j = (int)(idx - i * n_pairs);
    const unsigned long long row = first_row + (unsigned long long)i;
    uint32_t r[4];
    philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)j, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    out[i * ld + 2 * j] = (T)(rad * c);
    if (2 * j + 1 < dim) out[i * ld + 2 * j + 1] = (T)(rad * s);
}

}  // namespace jf
