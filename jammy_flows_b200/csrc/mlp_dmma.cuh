// fp64 parameter-generator MLP on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64) -- one hidden layer.
//
//   params[B, N] = W2 * tanh(W1 * in + b1) + b2        (reference main/default.py:654-670, hidden "128")
//
// Why DMMA and not tcgen05: tcgen05 has no f64 kind, and the parity contract is fp64 relative 1e-10.  Measured on
// this pool's B200 (tools/microbench.cu, profiles/microbench_r01.json): DFMA 33.8 TFLOP/s, DMMA m8n8k4 37.2 TFLOP/s,
// both together 31.9 TFLOP/s -- they share the FP64 pipe, so DMMA buys no peak, but it frees 8x the issue slots and
// most of the shared-memory operand traffic, which is what a CUDA-core fp64 GEMM runs out of.
//
// Layout: one warp owns 8 rows end to end (16 warps, 128 rows per CTA).
//   GEMM1   [8 x Kin] x [Kin x H]: A fragments from the gathered input tile in smem, B fragments from W1 in smem
//           (torch's [out,in] row-major weight IS the "col" operand layout of mma.m8n8k4).  Accumulators: H/8
//           C tiles in registers; bias + tanh applied in registers.
//   relayout the C-fragment layout (row, 2 cols per lane) is turned into the A-fragment layout (row, k = lane%4) with
//           four quad shuffles per 8x8 tile; the hidden activations never leave registers.
//   GEMM2   loop over N in tiles of 64 output columns: W2 tile [64 x H] double-buffered in smem with cp.async, shared
//           by the 16 warps; A comes from registers, so the main loop issues one LDS.64 per DMMA.
//   store   param-major (out[n*ld + row]) so the consuming layer kernel reads coalesced.
#pragma once
#include "common.cuh"
#include "mlp_kernels.cuh"

namespace jf {

JF_DEVINL void dmma884(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

JF_DEVINL void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
JF_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> JF_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

constexpr int kDmmaRows = 128;   // rows per CTA (16 warps x 8 rows)
constexpr int kDmmaThreads = 512;
constexpr int kDmmaTN = 64;      // W2 tile: output columns per smem stage

// leading dimension (in doubles) that makes the 8B fragment loads bank-conflict free: ld % 8 == 4
__host__ __device__ inline int dmma_ld(int k) { int kp = (k + 3) / 4 * 4; return (kp % 8 == 4) ? kp : kp + 4; }

// tanh is inlined ~100 instructions long; 2*HP/8 copies of it blow the instruction cache, so call it
__device__ __noinline__ double tanh_call(double x) { return tanh(x); }

// Each warp owns 8 rows (one m8 fragment row block).  With the hidden activations of 8 rows x HP in registers as
// A fragments (HP/4 doubles per lane) a thread needs ~110 registers, so 16 warps fit one CTA per SM: 4 warps per
// scheduler hide the DMMA / LDS latencies, and every warp has 4 independent accumulator chains per k-step.
template <int HP>   // hidden width padded to a multiple of 8 (<= 128)
__global__ void __launch_bounds__(kDmmaThreads, 1) mlp2_dmma_kernel(const __grid_constant__ MlpArgs<double> m) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    const int Kin = m.dims[0], H = m.dims[1], N = m.dims[2];
    const int ldin = dmma_ld(Kin);
    const int kin4 = (Kin + 3) / 4;
    constexpr int ldh = HP + 4;
    constexpr int NT1 = HP / 8;                       // hidden n-tiles
    constexpr int kStage = kDmmaTN * ldh + kDmmaTN;   // W2 tile + its bias slice (doubles)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;            // fragment row / quad index
    const int64_t row0 = (int64_t)blockIdx.x * kDmmaRows;

    // ---- stage 0: gather the input tile and W1 into smem (zero padded) ------------------------------------------
    double* sIn = smem;                                // [128][ldin]
    double* sW1 = smem + (size_t)kDmmaRows * ldin;     // [HP][ldin]
    for (int e = tid; e < kDmmaRows * ldin; e += kDmmaThreads) {
        const int r = e / ldin, c = e - r * ldin;
        const int64_t row = row0 + r;
        double v = 0.0;
        if (row < m.B && c < Kin) {
            int cc = c, s = 0;
            while (cc >= m.seg_cols[s]) { cc -= m.seg_cols[s]; ++s; }
            v = m.seg_ptr[s][row * m.seg_ld[s] + cc];
        }
        sIn[e] = v;
    }
    for (int e = tid; e < HP * ldin; e += kDmmaThreads) {
        const int n = e / ldin, c = e - n * ldin;
        sW1[e] = (n < H && c < Kin) ? m.wt[0][(size_t)n * Kin + c] : 0.0;
    }
    __syncthreads();

    // ---- GEMM1: hidden pre-activations of this warp's 8 rows ----------------------------------------------------
    double c1[NT1][2];
#pragma unroll
    for (int t = 0; t < NT1; ++t) {
        const int col = t * 8 + 2 * q;
        c1[t][0] = (col < H) ? m.bias[0][col] : 0.0;
        c1[t][1] = (col + 1 < H) ? m.bias[0][col + 1] : 0.0;
    }
    {
        const double* ap = sIn + (size_t)(warp * 8 + g) * ldin + q;
        const double* bp = sW1 + (size_t)g * ldin + q;
        for (int ks = 0; ks < kin4; ++ks) {
            const double a = ap[ks * 4];
#pragma unroll
            for (int t = 0; t < NT1; ++t) dmma884(c1[t][0], c1[t][1], a, bp[(size_t)t * 8 * ldin + ks * 4]);
        }
    }
    // ---- tanh + C-fragment -> A-fragment relayout (registers only) ----------------------------------------------
    double a2[2 * NT1];
    {
        const int src_lo = (lane & ~3) | (q >> 1), src_hi = src_lo + 2;
        const bool odd = q & 1;
#pragma unroll
        for (int t = 0; t < NT1; ++t) {
            const double v0 = tanh_call(c1[t][0]), v1 = tanh_call(c1[t][1]);
            const double s0 = __shfl_sync(0xffffffffu, v0, src_lo), s1 = __shfl_sync(0xffffffffu, v1, src_lo);
            const double t0 = __shfl_sync(0xffffffffu, v0, src_hi), t1 = __shfl_sync(0xffffffffu, v1, src_hi);
            a2[2 * t] = odd ? s1 : s0;
            a2[2 * t + 1] = odd ? t1 : t0;
        }
        // padded hidden units (col >= H) carry tanh(0) = 0 (zero weights and bias)
    }
    __syncthreads();   // everyone is done with sIn / sW1: the region is reused for the W2 stages

    // ---- GEMM2: stream W2 (+ its bias slice) in tiles of 64 output columns ---------------------------------------
    const double* __restrict__ W2 = m.wt[1];
    const double* __restrict__ B2 = m.bias[1];
    const int n_tiles = (N + kDmmaTN - 1) / kDmmaTN;
    auto load_tile = [&](int tile, int buf) {
        double* dst = smem + (size_t)buf * kStage;
        const int n0 = tile * kDmmaTN;
        for (int e = tid; e < kDmmaTN * (HP / 2); e += kDmmaThreads) {      // 16-byte chunks
            const int nn = e / (HP / 2), kc = (e - nn * (HP / 2)) * 2;
            const int n = n0 + nn;
            int bytes = 0;
            if (n < N) bytes = (kc + 1 < H) ? 16 : ((kc < H) ? 8 : 0);
            const double* src = (bytes > 0) ? (W2 + (size_t)n * H + kc) : W2;   // bytes == 0: pure zero fill
            cp_async16(dst + (size_t)nn * ldh + kc, src, bytes);
        }
        if (tid < kDmmaTN / 2) {                                              // bias slice, 2 doubles per thread
            const int n = n0 + 2 * tid;
            const int bytes = (n + 1 < N) ? 16 : ((n < N) ? 8 : 0);
            cp_async16(dst + (size_t)kDmmaTN * ldh + 2 * tid, bytes > 0 ? (B2 + n) : B2, bytes);
        }
        cp_async_commit();
    };
    load_tile(0, 0);
    const int64_t row = row0 + warp * 8 + g;
    for (int tile = 0; tile < n_tiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < n_tiles) { load_tile(tile + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* sW = smem + (size_t)buf * kStage;
        const double* sB = sW + (size_t)kDmmaTN * ldh;
        const int n0 = tile * kDmmaTN;
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            if (n0 + sub * 32 >= N) break;
            double acc[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[j][0] = sB[sub * 32 + j * 8 + 2 * q];
                acc[j][1] = sB[sub * 32 + j * 8 + 2 * q + 1];
            }
            const double* bp = sW + (size_t)(sub * 32 + g) * ldh + q;
#pragma unroll
            for (int ks = 0; ks < HP / 4; ++ks) {
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[j][0], acc[j][1], a2[ks], bp[(size_t)j * 8 * ldh + ks * 4]);
            }
            if (row < m.B) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = n0 + sub * 32 + j * 8 + 2 * q;
                    if (col < N) m.out[(int64_t)col * m.so_p + row * m.so_r] = acc[j][0];
                    if (col + 1 < N) m.out[(int64_t)(col + 1) * m.so_p + row * m.so_r] = acc[j][1];
                }
            }
        }
        __syncthreads();   // tile fully consumed before its buffer is refilled
    }
}

// shared memory the kernel needs for a given shape
inline size_t mlp2_dmma_smem(int Kin, int HP) {
    const size_t stage0 = (size_t)(kDmmaRows + HP) * dmma_ld(Kin);
    const size_t stage2 = (size_t)2 * (kDmmaTN * (HP + 4) + kDmmaTN);
    return (stage0 > stage2 ? stage0 : stage2) * sizeof(double);
}

}  // namespace jf
