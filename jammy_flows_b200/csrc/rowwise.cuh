// Linear layer with PER-ROW weights: the "being amortised" mode of the reference's AmortizableMLP
// (amortizable_mlp.py:470-585, `_adaptive_matmul` / `_apply_amortized_mlp` with use_permanent_parameters=False), which
// `pdf(..., amortize_everything=True)` / `fully_amortized_pdf` (main/fully_amortized.py:22-278) run for every sub-pdf:
//     out[r, o] (+)= act( sum_i W_r[o, i] * in[r, i] + b_r[o] ),  W_r = params[r, off_w + o*n_in + i],  b_r = params[r, off_b + o]
// Every weight is read exactly once, so the kernel is HBM-bound: (n_in*n_out + n_out) elements per row.
//
// One warp per row.  The row's weight block is contiguous in memory:
//   n_in >= 32: lanes stride over i for one output at a time (coalesced 256 B rows of W), shuffle reduction, results
//               of 32 consecutive outputs collected one per lane and finished (bias, tanh, store) coalesced;
//   n_in <  32: tiles of whole weight rows are staged in shared memory with coalesced cp.async copies (odd row stride:
//               no bank conflicts), then one lane per output walks its row.
#pragma once
#include "common.cuh"

namespace jf {

constexpr int ROWWISE_WARPS = 8;
constexpr int ROWWISE_TILE = 640;      // staged weight elements per warp (narrow case)

template <typename T>
__device__ __forceinline__ void rowwise_finish(T acc, int o, const T* __restrict__ p, int64_t off_b, int act,
                                               int accumulate, T* __restrict__ dst) {
    if (off_b >= 0) acc += p[off_b + o];
    if (act) acc = tanh(acc);
    if (accumulate) acc += *dst;
    *dst = acc;
}

template <typename T>
__global__ void __launch_bounds__(ROWWISE_WARPS * 32)
rowwise_linear_kernel(const T* __restrict__ params, int64_t ld_p, int64_t off_w, int64_t off_b,
                      const T* __restrict__ in, int64_t ld_in, int n_in, int n_in_pad, int n_out, int act,
                      int accumulate, T* __restrict__ out, int64_t so_p, int64_t so_r, int64_t R) {
    extern __shared__ __align__(16) unsigned char rowwise_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool wide = n_in >= 32;
    const int per_warp = n_in_pad + (wide ? 0 : ROWWISE_TILE);
    T* s_in = reinterpret_cast<T*>(rowwise_smem) + (size_t)warp * per_warp;
    T* s_w = s_in + n_in_pad;
    for (int64_t r = (int64_t)blockIdx.x * ROWWISE_WARPS + warp; r < R; r += (int64_t)gridDim.x * ROWWISE_WARPS) {
        const T* p = params + r * ld_p;
        const T* w = p + off_w;
        T* o_row = out + r * so_r;
        for (int i = lane; i < n_in; i += 32) s_in[i] = in[r * ld_in + i];
        __syncwarp();
        if (wide) {
            T res = T(0);
            constexpr int kUnrollO = sizeof(T) == 8 ? 2 : 1;      // measured: fp64 +1 %, fp32 -18 % with two outputs in flight
#pragma unroll kUnrollO
            for (int o = 0; o < n_out; ++o) {
                const T* wr = w + (int64_t)o * n_in;
                T a0 = T(0), a1 = T(0);
                int i = lane;
                for (; i + 32 < n_in; i += 64) {
                    a0 = fma(wr[i], s_in[i], a0);
                    a1 = fma(wr[i + 32], s_in[i + 32], a1);
                }
                if (i < n_in) a0 = fma(wr[i], s_in[i], a0);
                T acc = a0 + a1;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
                if (lane == (o & 31)) res = acc;
                if ((o & 31) == 31 || o == n_out - 1) {
                    const int oo = (o & ~31) + lane;
                    if (oo <= o) rowwise_finish(res, oo, p, off_b, act, accumulate, o_row + oo * so_p);
                }
            }
        } else {
            const int stride = n_in | 1;
            const int rows_per_tile = ROWWISE_TILE / stride;
            const int step_r = 32 / n_in, step_c = 32 - step_r * n_in;
            for (int o0 = 0; o0 < n_out; o0 += rows_per_tile) {
                const int nr = min(rows_per_tile, n_out - o0);
                const int n = nr * n_in;
                const T* wt = w + (int64_t)o0 * n_in;
                // element e = lane + 32 j of the tile lives at (row e / n_in, column e % n_in): both advance incrementally
                // global -> shared without a register round trip (cp.async, one element each: the block may start at
                // any element, so wider copies are not aligned): the whole tile is in flight before the first wait
                int rr = lane / n_in, cc = lane - rr * n_in;
                for (int e = lane; e < n; e += 32) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(s_w + rr * stride + cc);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(dst), "l"(wt + e), "n"(sizeof(T)));
                    rr += step_r;
                    cc += step_c;
                    if (cc >= n_in) { cc -= n_in; ++rr; }
                }
                asm volatile("cp.async.commit_group;\n" ::);
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                __syncwarp();
                for (int oo = lane; oo < nr; oo += 32) {
                    const T* wr = s_w + oo * stride;
                    T acc = T(0);
                    for (int i = 0; i < n_in; ++i) acc = fma(wr[i], s_in[i], acc);
                    rowwise_finish(acc, o0 + oo, p, off_b, act, accumulate, o_row + (int64_t)(o0 + oo) * so_p);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
}

}  // namespace jf
