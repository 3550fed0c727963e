// Backward of the parameter generator  params[P, B] = W2 tanh(W1 x + b1) + b2  (fp32, hidden 128) on the tensor cores.
//
//   reference: autograd through nn.Sequential(Linear, Tanh, Linear), main/default.py:654-670, called at :956; here the
//   upstream gradient G = d loss / d params arrives PARAM-MAJOR ([P, B], what csrc/gf_bwd.cuh writes).
//
// The two products that carry 99 % of the flops (2 x 2 P 128 B) run as tcgen05.mma kind::tf32 with fp32 accumulators in
// tensor memory -- tf32 keeps 10 mantissa bits of each operand: the gradients of the generator weights carry a relative
// error of ~1e-3 / sqrt(terms) (stated tolerance, tests/test_cuda_training.py: 2e-3 of the tensor maximum against fp64):
//   bw_dh_kernel    dh[B,128]   = G^T W2            M = 128 rows, N = 128, K = P   (one CTA per 128 rows)
//                   epilogue: dpre = dh (1 - h^2) s -> [B,128] row major   (s: optional per-row scale of the upstream gradient)
//   bw_dw2_kernel   [dW2 | db2] = G [h | 1]         M = 128 parameters, N = 144, K = rows (CTA = parameter tile x row range,
//                   partial sums added with red.global)
// Operands are staged in shared memory in the canonical K-major no-swizzle UMMA layout (8 x 16 B core matrices; the same
// layout as csrc/mlp_i8.cuh): G is read ONCE per kernel straight from HBM (cp.async; bw_dh transposes its chunk through
// shared memory, for bw_dw2 the param-major rows already are K-major), W2 and h come pre-tiled from small prep kernels
// (bw_w2_tiles_kernel once per call, bw_h_tiles_kernel: h = tanh(W1 x + b1) recomputed in fp32) so that one 1-D
// cp.async.bulk per chunk is enough.  Both kernels are HBM bound by design: 2 x P B 4 bytes of G per call.
// The three small products (dW1 = dpre^T x, db1, dx = dpre W1: 1 % of the flops) are a plain FFMA kernel (bw_small_kernel).
#pragma once
#include <cuda.h>
#include "mlp_i8.cuh"
#include "mlp_bwd_launch.cuh"

namespace jf {

constexpr int kBwH = 128;          // hidden width
constexpr int kBwKC = 32;          // K elements per pipeline chunk (4 MMAs of K = 8)
constexpr int kBwNExt = 144;       // [h | 1 | padding]: N of the dW2 product
constexpr int kBwProducers = 256;  // 8 warps stage operands, warp 8 issues the MMAs
constexpr int kBwThreads = 288;


// byte offset of element (r, k) of a [R rows][KC k] fp32 operand tile in the K-major no-swizzle UMMA layout
__host__ __device__ constexpr int bw_tile_off(int R, int r, int k) { return (k >> 2) * (R / 8) * 128 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4; }

JF_DEVINL float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
JF_DEVINL void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor of kind::tf32: D = f32, A = B = tf32, both K-major
__host__ __device__ constexpr uint32_t bw_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
JF_DEVINL void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
JF_DEVINL void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// ---- prep: W2 [P,128] -> tiles of 32 parameters: B operand (n = hidden, k = parameter) of the dh product ------------------
__global__ void __launch_bounds__(256) bw_w2_tiles_kernel(const float* __restrict__ W2, int P, float* __restrict__ tiles) {
    const int c = blockIdx.x;
    char* t = reinterpret_cast<char*>(tiles) + (size_t)c * (kBwH * kBwKC * 4);
    for (int e = threadIdx.x; e < kBwH * kBwKC; e += blockDim.x) {
        const int k = e >> 7, n = e & 127;                  // consecutive threads: consecutive n of one parameter (coalesced)
        const int p = c * kBwKC + k;
        const float v = p < P ? W2[(size_t)p * kBwH + n] : 0.f;
        *reinterpret_cast<float*>(t + bw_tile_off(kBwH, n, k)) = to_tf32(v);
    }
}

// ---- prep: h = tanh(W1 x + b1) -> tiles of 32 rows: B operand (n = hidden | 1, k = row) of the dW2 product ---------------
// persistent blocks over tiles (W1 is staged in shared memory once per block, not once per 32 rows); thread (row = tid & 31,
// group of 16 hidden units = tid >> 5)
__global__ void __launch_bounds__(256) bw_h_tiles_kernel(const BwArgs a, int64_t n_tiles) {
    extern __shared__ float sm_h[];
    float* sW = sm_h;                                       // [in][128]  (W1 transposed)
    float* sX = sW + (size_t)a.in * kBwH;                   // [32][in + 1]
    const int ldxs = a.in | 1;
    for (int e = threadIdx.x; e < a.in * kBwH; e += blockDim.x) {
        const int u = e / a.in, i = e - u * a.in;
        sW[i * kBwH + u] = a.W1[e];
    }
    const int r = threadIdx.x & 31, ug = threadIdx.x >> 5;
    float bias[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bias[j] = a.b1[ug * 16 + j];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kBwKC;
        __syncthreads();                                    // (first pass: W1 is in place; later: the previous tile's inputs are consumed)
        for (int e = threadIdx.x; e < kBwKC * a.in; e += blockDim.x) {
            const int rr = e / a.in, i = e - rr * a.in;
            sX[rr * ldxs + i] = (row0 + rr < a.B) ? a.x[(row0 + rr) * a.ldx + i] : 0.f;
        }
        __syncthreads();
        const bool live = row0 + r < a.B;
        char* t = reinterpret_cast<char*>(a.h_tiles) + (size_t)tile * (kBwNExt * kBwKC * 4);
        float z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = bias[j];
        for (int i = 0; i < a.in; ++i) {
            const float xi = sX[r * ldxs + i];
            const float* w = sW + i * kBwH + ug * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = fmaf(xi, w[j], z[j]);
        }
        // the upstream gradient of row r is G[., r] * s_r (s = 1 without row_scale): the scale rides on the B operand of the
        // dW2 product (h s | s) and on the factor (1 - h^2) s of the dh epilogue, G itself is never rescaled
        const float sr = (live && a.row_scale != nullptr) ? a.row_scale[row0 + r] : 1.f;
        float fc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float h = tanhf(z[j]);
            *reinterpret_cast<float*>(t + bw_tile_off(kBwNExt, ug * 16 + j, r)) = live ? to_tf32(h * sr) : 0.f;
            fc[j] = (1.f - h * h) * sr;
        }
        if (live) {
            float4* dst = reinterpret_cast<float4*>(a.fac + (row0 + r) * kBwH + ug * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_float4(fc[4 * q], fc[4 * q + 1], fc[4 * q + 2], fc[4 * q + 3]);
        }
        // column 128 = s (its product with G is db2), the padding columns are zero
        if (ug < 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = kBwH + ug * 8 + j;
                *reinterpret_cast<float*>(t + bw_tile_off(kBwNExt, n, r)) = (n == kBwH && live) ? to_tf32(sr) : 0.f;
            }
        }
    }
}

// ---- dh = G^T W2, dpre = dh (1 - h^2) ---------------------------------------------------------------------------------------
// A = G^T: M = 128 rows, K = parameters.  G is param-major (rows contiguous), i.e. the A operand is MN-MAJOR: TMA boxes of
// 32 rows x 32 parameters in the 128-byte swizzle with 32-byte atoms ARE the canonical MN-major tf32 UMMA layout (32 fp32 of M
// per 128-byte line, one line per K index, 4-line groups 512 bytes apart = SBO, the four row boxes of a chunk 4096 bytes
// apart = LBO), so the
// gradient block goes from HBM to the tensor core without passing through registers (the first version transposed every
// chunk through shared memory with 8 producer warps).  3 stages of (A 16 KB, W2 tile 16 KB); 2 CTAs per SM.
constexpr int kDhStages = 3;
constexpr int kDhTile = kBwH * kBwKC * 4;                   // 16 KB: A chunk (4 boxes of 4 KB) and W2 tile alike
constexpr int kDhBox = 32 * 32 * 4;
constexpr int kDhSmem = 2 * kDhStages * kDhTile + 256 + 1024;

// MN-major tf32 operands have exactly one shared-memory layout: the 128-byte swizzle with a 32-byte base (layout type 1,
// TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32 fp32 of M per 128-byte line, one line per K index, the 32-byte pieces of a
// line permuted by (line & 3); lbo = distance between 32-element chunks of M, sbo = between groups of 4 lines (4 K)
JF_DEVINL uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)1 << 61);
}
JF_DEVINL void tma_load_2d_(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(kBwThreads, 2) bw_dh_kernel(const BwArgs a, const __grid_constant__ CUtensorMap tmapG) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sraw = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (sbase - sraw);
    const uint32_t offB = kDhStages * kDhTile, offBar = 2 * kDhStages * kDhTile;
    const uint32_t bar0 = sbase + offBar;
    auto bar_full = [&](int s) { return bar0 + 8 * s; };
    auto bar_empty = [&](int s) { return bar0 + 8 * (4 + s); };
    const uint32_t bar_done = bar0 + 8 * 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + offBar + 128);
    const int64_t row0 = (int64_t)blockIdx.x * 128;
    const int n_chunks = (a.P + kBwKC - 1) / kBwKC;

    if (tid == 0) {
        for (int s = 0; s < kDhStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar0 + 128), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- producer: one thread; per chunk of 32 parameters four TMA boxes (32 rows each) of G and the W2 tile ----
        if (elect_one()) {
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % kDhStages;
                if (c >= kDhStages) mbar_wait(bar_empty(s), (uint32_t)(((c / kDhStages) - 1) & 1));
                mbar_expect_tx(bar_full(s), 2 * kDhTile);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    tma_load_2d_(sbase + s * kDhTile + b * kDhBox, &tmapG, (int)row0 + 32 * b, c * kBwKC, bar_full(s));
                bulk_g2s(sbase + offB + s * kDhTile, a.w2_tiles + (size_t)c * (kDhTile / 4), kDhTile, bar_full(s));
            }
        }
    } else if (warp == 8) {
        // ---- MMA issuer ----
        if (elect_one()) {
            constexpr uint32_t idesc = bw_idesc(128, 128) | (1u << 15);              // A is MN-major
            constexpr uint32_t lboB = (128 / 8) * 128;
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % kDhStages;
                mbar_wait(bar_full(s), (uint32_t)((c / kDhStages) & 1));
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t ad = umma_desc_mn_sw128(sbase + s * kDhTile + ks * 1024, kDhBox, 512);
                    const uint64_t bd = umma_desc(sbase + offB + s * kDhTile + ks * 2 * lboB, lboB, 128);
                    tc_mma_tf32(tmem, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                }
                tc_commit(bar_empty(s));
            }
            tc_commit(bar_done);
        }
    }
    // ---- epilogue: warps 0-3, thread = row ----
    if (warp < 4) {
        __syncwarp();
        mbar_wait(bar_done, 0);
        tc_fence_after();
        const int r = warp * 32 + lane;
        const int64_t row = row0 + r;
        const float* fc = a.fac + row * kBwH;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int n0 = 0; n0 < kBwH; n0 += 8) {
            int v[8];
            tmem_ld8(tl + n0, v);
            tmem_ld_wait();
            if (row < a.B) {
                const float4 f0 = *reinterpret_cast<const float4*>(fc + n0), f1 = *reinterpret_cast<const float4*>(fc + n0 + 4);
                const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = __int_as_float(v[j]) * f[j];
                float4* dst = reinterpret_cast<float4*>(a.dpre + row * kBwH + n0);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
}

// ---- [dW2 | db2] = G [h | 1] ------------------------------------------------------------------------------------------------
// CTA = (parameter tile of 128, row range); 3 stages of (A: 128 p x 64 rows of G, 32 KB; B: two h tiles of 144 x 32, 36 KB).
// The A operand is loaded by TMA (cp.async.bulk.tensor.2d, two boxes of 32 rows x 128 parameters per chunk) straight into
// the 128-byte-swizzled K-major layout that tcgen05.mma reads: one elected thread issues the copies, there are no producer
// warps and no per-thread cp.async (the 16-byte cp.async version kept the l1tex pipe 59 % busy at 28 % of the DRAM rate,
// long_scoreboard 11.7 per issue -- ncu, profiles/ncu_r02_train.md); rows >= B and parameters >= P are zero-filled by the
// tensor map.  64 rows per chunk = 256 contiguous bytes of every parameter row.
constexpr int kW2Stages = 3;
constexpr int kW2KC = 64;                                    // rows per chunk (8 MMAs of K = 8)
constexpr int kW2BoxA = 128 * 32 * 4;                        // one TMA box: 128 parameters x 32 rows (128 bytes per row: the swizzle atom)
constexpr int kW2TileA = 2 * kW2BoxA, kW2TileB = kBwNExt * kBwKC * 4;     // B stage = 2 h tiles
constexpr int kW2Smem = kW2Stages * (kW2TileA + 2 * kW2TileB) + 256 + 1024;   // (+ alignment of the swizzled stages to 1024 bytes)

// K-major operand in the 128-byte swizzle: rows of 128 bytes (32 fp32 of K), 8-row groups 1024 bytes apart
JF_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
JF_DEVINL void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(kBwThreads, 1) bw_dw2_kernel(const BwArgs a, const __grid_constant__ CUtensorMap tmapG) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sraw = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t sbase = (sraw + 1023u) & ~1023u;                                    // swizzle atoms are 1024-byte aligned
    unsigned char* smem = smem_raw + (sbase - sraw);
    const uint32_t offB = kW2Stages * kW2TileA, offBar = offB + kW2Stages * 2 * kW2TileB;
    const uint32_t bar0 = sbase + offBar;
    auto bar_full = [&](int s) { return bar0 + 8 * s; };
    auto bar_empty = [&](int s) { return bar0 + 8 * (4 + s); };
    const uint32_t bar_done = bar0 + 8 * 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + offBar + 128);
    const int p0 = blockIdx.x * 128;
    // this CTA's row range, in chunks of 64 rows
    const int64_t total_chunks = (a.B + kW2KC - 1) / kW2KC;
    const int64_t per = (total_chunks + a.n_splits - 1) / a.n_splits;
    const int64_t c_begin = (int64_t)blockIdx.y * per;
    const int64_t c_end = c_begin + per < total_chunks ? c_begin + per : total_chunks;
    const int n_chunks = c_end > c_begin ? (int)(c_end - c_begin) : 0;

    if (tid == 0) {
        for (int s = 0; s < kW2Stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar0 + 128), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (n_chunks > 0) {
        if (warp == 0) {
            // ---- producer: one thread, TMA for G and a bulk copy for the two h tiles of every chunk ----
            if (elect_one()) {
                for (int c = 0; c < n_chunks; ++c) {
                    const int s = c % kW2Stages;
                    if (c >= kW2Stages) mbar_wait(bar_empty(s), (uint32_t)(((c / kW2Stages) - 1) & 1));
                    const int64_t row0 = (c_begin + c) * kW2KC;
                    mbar_expect_tx(bar_full(s), kW2TileA + 2 * kW2TileB);
                    tma_load_2d(sbase + s * kW2TileA, &tmapG, (int)row0, p0, bar_full(s));
                    tma_load_2d(sbase + s * kW2TileA + kW2BoxA, &tmapG, (int)row0 + 32, p0, bar_full(s));
                    bulk_g2s(sbase + offB + s * 2 * kW2TileB, a.h_tiles + (size_t)(c_begin + c) * (2 * kW2TileB / 4), 2 * kW2TileB,
                             bar_full(s));
                }
            }
        } else if (warp == 8) {
            if (elect_one()) {
                constexpr uint32_t idesc = bw_idesc(128, kBwNExt);
                constexpr uint32_t lboB = (kBwNExt / 8) * 128;
                for (int c = 0; c < n_chunks; ++c) {
                    const int s = c % kW2Stages;
                    mbar_wait(bar_full(s), (uint32_t)((c / kW2Stages) & 1));
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < kW2KC / 8; ++ks) {
                        // A: box (ks >> 2), 32 bytes (8 fp32 of K) further inside the swizzle atom per step
                        const uint64_t ad = umma_desc_sw128(sbase + s * kW2TileA + (ks >> 2) * kW2BoxA + (ks & 3) * 32);
                        const uint64_t bd = umma_desc(sbase + offB + s * 2 * kW2TileB + (ks >> 2) * kW2TileB + (ks & 3) * 2 * lboB, lboB, 128);
                        tc_mma_tf32(tmem, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                    }
                    tc_commit(bar_empty(s));
                }
                tc_commit(bar_done);
            }
        }
        // ---- epilogue: warps 0-3, thread = parameter row: partial sums -> red.global ----
        if (warp < 4) {
            __syncwarp();
            mbar_wait(bar_done, 0);
            tc_fence_after();
            const int p = p0 + warp * 32 + lane;
            const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int n0 = 0; n0 < kBwNExt - 8; n0 += 8) {                       // columns 0..135 (128 = db2, the rest is padding)
                int v[8];
                tmem_ld8(tl + n0, v);
                tmem_ld_wait();
                if (p < a.P) {
                    if (n0 < kBwH) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) red_add(a.dW2 + (size_t)p * kBwH + n0 + j, __int_as_float(v[j]));
                    } else {
                        red_add(a.db2 + p, __int_as_float(v[0]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

// ---- dW1 = dpre^T x, db1 = sum dpre, dx = dpre W1 (1 % of the flops: FFMA) -------------------------------------------------
// persistent blocks over tiles of 64 rows.  dW1: thread -> (pair of hidden units 2 hp, 2 hp + 1; quarter q of the inputs):
// per row one 8-byte load of the two dpre values and 16-byte broadcast loads of the inputs feed 2 x kSmPer multiply-adds
// (the first version issued 9 shared-memory loads per 8 multiply-adds).  The partial sums stay in registers over all
// tiles of a block and go out with ONE red.global pass per block.
constexpr int kSmRows = 64;
constexpr int kSmPer = 24;                              // inputs per thread: ceil(96 / 4)
constexpr int kSmLdD = 130;                             // dpre row stride in shared memory (even: 8-byte loads, conflict free)
__global__ void __launch_bounds__(256) bw_small_kernel(const BwArgs a) {
    extern __shared__ __align__(16) float sm_s[];
    const int ldxs = ((a.in + 3) / 4) * 4 + 4 * ((kSmPer + 3) / 4);   // inputs padded so that every thread may read kSmPer values
    float* sD = sm_s;                                   // [64][130] dpre
    float* sX = sD + kSmRows * kSmLdD;                  // [64][ldxs] inputs (zero padded)
    float* sW = sX + kSmRows * ldxs;                    // [128][in | 1]  W1 (dx only)
    const int ldw = a.in | 1;
    if (a.dx != nullptr)
        for (int e = threadIdx.x; e < kBwH * a.in; e += blockDim.x) {
            const int u = e / a.in, i = e - u * a.in;
            sW[u * ldw + i] = a.W1[e];
        }
    const int hp = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int per = ((a.in + 3) / 4 + 3) / 4 * 4;      // inputs per quarter, a multiple of 4 (<= kSmPer)
    const int i0 = q * per;
    float acc0[kSmPer], acc1[kSmPer];
#pragma unroll
    for (int j = 0; j < kSmPer; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
    float bsum0 = 0.f, bsum1 = 0.f;
    const int64_t n_tiles = (a.B + kSmRows - 1) / kSmRows;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kSmRows;
        __syncthreads();                                // the previous tile has been consumed
        for (int e = threadIdx.x; e < kSmRows * kBwH; e += blockDim.x) {
            const int r = e >> 7, n = e & 127;
            sD[r * kSmLdD + n] = (row0 + r < a.B) ? a.dpre[(row0 + r) * kBwH + n] : 0.f;
        }
        for (int e = threadIdx.x; e < kSmRows * ldxs; e += blockDim.x) {
            const int r = e / ldxs, i = e - r * ldxs;
            sX[e] = (i < a.in && row0 + r < a.B) ? a.x[(row0 + r) * a.ldx + i] : 0.f;
        }
        __syncthreads();
        // dW1[u][i], db1[u]
#pragma unroll 2
        for (int r = 0; r < kSmRows; ++r) {
            const float2 dv = *reinterpret_cast<const float2*>(sD + r * kSmLdD + 2 * hp);
            const float4* xr = reinterpret_cast<const float4*>(sX + r * ldxs + i0);
            if (q == 0) { bsum0 += dv.x; bsum1 += dv.y; }
#pragma unroll
            for (int j4 = 0; j4 < kSmPer / 4; ++j4) {
                if (4 * j4 < per) {
                    const float4 xv = xr[j4];
                    acc0[4 * j4 + 0] = fmaf(dv.x, xv.x, acc0[4 * j4 + 0]); acc1[4 * j4 + 0] = fmaf(dv.y, xv.x, acc1[4 * j4 + 0]);
                    acc0[4 * j4 + 1] = fmaf(dv.x, xv.y, acc0[4 * j4 + 1]); acc1[4 * j4 + 1] = fmaf(dv.y, xv.y, acc1[4 * j4 + 1]);
                    acc0[4 * j4 + 2] = fmaf(dv.x, xv.z, acc0[4 * j4 + 2]); acc1[4 * j4 + 2] = fmaf(dv.y, xv.z, acc1[4 * j4 + 2]);
                    acc0[4 * j4 + 3] = fmaf(dv.x, xv.w, acc0[4 * j4 + 3]); acc1[4 * j4 + 3] = fmaf(dv.y, xv.w, acc1[4 * j4 + 3]);
                }
            }
        }
        // dx[row][i] = sum_u dpre[row][u] W1[u][i]: thread -> (row = t & 63, inputs i = (t >> 6) + 4 j)
        if (a.dx != nullptr) {
            const int r = threadIdx.x & 63;
            if (row0 + r < a.B) {
                for (int i = threadIdx.x >> 6; i < a.in; i += 4) {
                    float s_ = 0.f;
                    for (int n = 0; n < kBwH; ++n) s_ = fmaf(sD[r * kSmLdD + n], sW[n * ldw + i], s_);
                    a.dx[(row0 + r) * a.lddx + i] = s_;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kSmPer; ++j) {
        const int i = i0 + j;
        if (j < per && i < a.in) {
            red_add(a.dW1 + (size_t)(2 * hp) * a.in + i, acc0[j]);
            red_add(a.dW1 + (size_t)(2 * hp + 1) * a.in + i, acc1[j]);
        }
    }
    if (q == 0) { red_add(a.db1 + 2 * hp, bsum0); red_add(a.db1 + 2 * hp + 1, bsum1); }
}

}  // namespace jf
