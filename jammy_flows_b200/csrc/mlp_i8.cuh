// fp64 parameter-generator MLP on the 5th-generation tensor cores (tcgen05, TMEM accumulators), one hidden layer of 128:
//
//   params[B, N] = W2 * tanh(W1 * in + b1) + b2        (reference main/default.py:654-670, hidden "128")
//
// tcgen05 has no f64 kind, and the parity contract is fp64 (1e-10 on log_pdf), so the dominant contraction (128 x N,
// 97 % of the flops) is computed EXACTLY in integer arithmetic (an Ozaki-style split):
//   * h = tanh(.) in [-1,1] and W2/scale_j (scale_j = power of two >= max_k |W2[j,k]|) are rounded to fixed point with
//     FA = 8*NS-2 fractional bits and written as NS balanced base-256 digits d in [-128,127]  (int8 "slices");
//   * slice p of h times slice q of W2 is an int8 x int8 -> int32 GEMM on tcgen05 (kind::i8): products are < 2^14, a
//     K = 128 dot product < 2^21, and all pairs with the same weight 256^-(p+q) share one TMEM accumulator (< 2^24):
//     no rounding anywhere in the tensor-core part;
//   * pairs with p+q >= NS (weight <= 2^-8NS relative) are dropped: for NS = 7 the result is within 2e-15 of the exact
//     product (tools/ozaki_probe.py), i.e. as good as an fp64 FMA chain; NS*(NS+1)/2 = 28 MMAs of 128x64x128;
//   * the epilogue reads the NS level accumulators back (tcgen05.ld), combines them by Horner in fp64, multiplies by
//     scale_j*2^-12 and adds b2: 2 fp64 operations per level instead of 256 per output in an fp64 FMA GEMM.
// The first layer (Kin <= 16 inputs) and tanh stay fp64 on the CUDA cores (3 % of the flops).
//
// One CTA owns 128 rows (UMMA M = 128, cta_group::1):
//   prologue  512 threads: gather inputs, layer 1 + tanh, digits -> A slices in shared memory in the canonical
//             K-major no-swizzle UMMA layout (8x16B core matrices), fence.proxy.async
//   main      warp 0 / lane 0  issues tcgen05.mma, tcgen05.commit -> mbarriers
//             warp 1 / lane 0  streams the pre-sliced W2 tiles (L2 resident, 8 KB per slice) with cp.async.bulk
//                              (the TMA engine, plain 1-D copies: the tiles are stored pre-tiled by the prep kernel)
//             warps 4-15       epilogue: TMEM -> registers -> fp64 -> param-major global stores
#pragma once
#include "common.cuh"
#include "mlp_kernels.cuh"

namespace jf {

constexpr int kI8Rows = 128;      // rows per CTA = UMMA M
constexpr int kI8H = 128;         // hidden width = K (bytes per slice row)
constexpr int kI8Threads = 512;    // 16 warps: 0 MMA issue, 1 W2 producer, 2 TMEM alloc, 4-15 epilogue; all 16 in the prologue
constexpr int kI8EpiWarps = 12;
constexpr int kI8MaxKin = 16;     // fp64: inputs of a row live in registers (kernel instantiated for 8 and 16: -17 % prologue time at 8)
constexpr int kI8MaxKinF32 = 96;  // fp32: inputs are streamed from shared memory
constexpr int kI8MaxSlots = 12;   // ring of W2-slice buffers (as many as fit next to the A slices, >= NS + 1)

template <int NS, int TN>
struct I8Cfg {
    static constexpr int kAccStages = (2 * NS * TN <= 512) ? 2 : 1;
    static constexpr int kTmemCols = 512;
    static constexpr int kSliceBytesA = kI8Rows * kI8H;       // 16 KB
    static constexpr int kSliceBytesB = TN * kI8H;            // 8 KB for TN = 64
    static constexpr int kFA = 8 * NS - 2;                    // fractional bits of the fixed-point operands
    static constexpr int kLboA = (kI8Rows / 8) * 128, kLboB = (TN / 8) * 128, kSbo = 128;
    // smem carve-up (bytes)
    static constexpr int offA = 0;
    static constexpr int offB = offA + NS * kSliceBytesA;
    // after the n_slots W2-slice buffers: 512 B of mbarriers + tmem pointer, 4*TN doubles of per-tile constants,
    // W1^T [Kin][128] + b1[128] doubles, the gathered inputs [128][Kin|1]
    __host__ __device__ static constexpr int tail_bytes(int kin, int es) { return 512 + 4 * TN * 8 + (kin + 1) * kI8H * es + kI8Rows * (kin | 1) * es; }
    __host__ __device__ static constexpr int smem_bytes(int kin, int n_slots, int es) { return offB + n_slots * kSliceBytesB + tail_bytes(kin, es); }
};

// bytes of device workspace for the pre-sliced W2 of an [N, 128] last layer: tiles + per-column scale (double)
template <int NS, int TN>
__host__ __device__ inline int64_t i8_prep_bytes(int N) {
    const int64_t n_tiles = (N + TN - 1) / TN;
    return n_tiles * NS * TN * kI8H + n_tiles * TN * 8;
}

// ---- PTX wrappers -----------------------------------------------------------------------------------------------------
JF_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// The suspend-time hint keeps a waiting warp asleep inside the instruction (it wakes when the phase completes): without
// it the try_wait / branch loops of the waiting warps were 30 % of all executed instructions of the fused kernel, and
// these kernels are instruction-issue bound (ncu, profiles/).
#ifndef JF_MBAR_SUSPEND_NS
#define JF_MBAR_SUSPEND_NS 20000
#endif
// JF_MBAR_SLEEP_NS > 0: back off with nanosleep after a failed try (a spinning warp otherwise issues try_wait / branch /
// yield continuously: a third of the executed instructions of the fused sampling kernel, ncu source page)
#ifndef JF_MBAR_SLEEP_NS
#define JF_MBAR_SLEEP_NS 0
#endif
JF_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"((uint32_t)JF_MBAR_SUSPEND_NS) : "memory");
        if (JF_MBAR_SLEEP_NS > 0 && !ok) asm volatile("nanosleep.u32 %0;" ::"r"((uint32_t)JF_MBAR_SLEEP_NS));
    } while (!ok);
}
JF_DEVINL void mbar_arrive(uint32_t bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
JF_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
JF_DEVINL void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
JF_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
JF_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
JF_DEVINL void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M = 128, K = 32 per instruction
JF_DEVINL void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: 8-row x 16-byte core matrices; lbo = byte distance between core matrices adjacent in K,
// sbo = between 8-row groups (cute::UMMA::SmemDescriptor, version 1)
JF_DEVINL uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// the four K=32 steps of one (slice p, slice q) pair in one go: descriptors advance by adding to their low words
// (start-address field, 16-byte units); a_step/b_step = 2*LBO >> 4
JF_DEVINL void tc_mma_i8_x4(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi_a, uint32_t desc_hi_b,
                            uint32_t idesc, uint32_t acc_first, uint32_t a_step, uint32_t b_step) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\n.reg .b32 al, bl;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "mov.b64 da, {%1, %3};\nmov.b64 db, {%2, %4};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n"
        "setp.ne.b32 p, 1, 0;\n"
        "add.u32 al, %1, %7;\nadd.u32 bl, %2, %8;\nmov.b64 da, {al, %3};\nmov.b64 db, {bl, %4};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n"
        "add.u32 al, al, %7;\nadd.u32 bl, bl, %8;\nmov.b64 da, {al, %3};\nmov.b64 db, {bl, %4};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n"
        "add.u32 al, al, %7;\nadd.u32 bl, bl, %8;\nmov.b64 da, {al, %3};\nmov.b64 db, {bl, %4};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi_a), "r"(desc_hi_b), "r"(idesc), "r"(acc_first), "r"(a_step),
          "r"(b_step) : "memory");
}
JF_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
JF_DEVINL void tmem_ld8(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
JF_DEVINL void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
JF_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tanh with an ABSOLUTE error of ~2 ulp(1): (1-e)/(1+e), e = exp(-2|x|); what the fixed-point operand needs
JF_DEVINL double tanh_abs(double x) {
    const double e = exp_neg(-2.0 * fabs(x));
    const double t = (1.0 - e) * rcp_1to2(1.0 + e);
    return x < 0.0 ? -t : t;
}

JF_DEVINL float tanh_abs(float x) {
    const float e = expf(-2.0f * fabsf(x));
    const float t = (1.0f - e) / (1.0f + e);
    return x < 0.0f ? -t : t;
}

// balanced base-256 digits of the fixed-point value round(v * 2^FA): byte s of the result is digit s (two's complement
// int8), s = 0 least significant.  v + C makes every digit an unsigned byte (carries included), ^C recentres it.
template <int NS>
JF_DEVINL unsigned long long to_digits(double v) {
    constexpr unsigned long long C = 0x8080808080808080ull >> (8 * (8 - NS));
    const long long q = __double2ll_rn(v * (double)(1ull << (8 * NS - 2)));
    return ((unsigned long long)q + C) ^ C;
}
template <int NS>
JF_DEVINL unsigned long long to_digits(float v) {      // NS <= 4: 32-bit fixed point (|v| <= 1 -> |q| <= 2^(8NS-2))
    static_assert(NS <= 4, "fp32 operands carry 24 bits: at most 4 slices");
    constexpr unsigned C = 0x80808080u >> (8 * (4 - NS));
    const int q = __float2int_rn(v * (float)(1u << (8 * NS - 2)));
    return (unsigned long long)(((unsigned)q + C) ^ C);
}

// ---- prep: W2 [N,128] fp64 -> NS int8 slices per tile of TN output columns, in the UMMA smem layout -------------------
// ws layout: [tile][slice q][TN*128 bytes], then double scale[n_tiles*TN] (= 2^e_j * 2^-12; 0 for padded columns)
template <typename T, int NS, int TN>
__global__ void __launch_bounds__(128) mlp_i8_prep_kernel(const T* __restrict__ W2, int N, unsigned char* ws) {
    const int tile = blockIdx.x, n_tiles = gridDim.x;
    double* scl = reinterpret_cast<double*>(ws + (size_t)n_tiles * NS * TN * kI8H);
    __shared__ double s_inv[TN];
    for (int jj = threadIdx.x; jj < TN; jj += blockDim.x) {
        const int j = tile * TN + jj;
        double mx = 0.0;
        if (j < N)
            for (int k = 0; k < kI8H; ++k) mx = fmax(mx, fabs((double)W2[(size_t)j * kI8H + k]));
        int e = 0;
        if (mx > 0.0) { frexp(mx, &e); }                 // mx = f * 2^e, f in [0.5,1)  =>  |W/2^e| < 1
        s_inv[jj] = ldexp(1.0, -e);
        // fp32 callers read the (exact, power-of-two) scale as a float from the low half of the 8-byte cell: no per-element
        // double -> float conversion in the epilogue
        if (sizeof(T) == 8) scl[tile * TN + jj] = (j < N) ? ldexp(1.0, e - 12) : 0.0;
        else { float* sf = reinterpret_cast<float*>(scl + tile * TN + jj); sf[0] = (j < N) ? ldexpf(1.0f, e - 12) : 0.0f; sf[1] = 0.0f; }
    }
    __syncthreads();
    unsigned char* base = ws + (size_t)tile * NS * TN * kI8H;
    for (int idx = threadIdx.x; idx < TN * kI8H; idx += blockDim.x) {
        const int jj = idx / kI8H, k = idx - jj * kI8H;
        const int j = tile * TN + jj;
        const T w = (j < N) ? (T)((double)W2[(size_t)j * kI8H + k] * s_inv[jj]) : T(0);   // power-of-two scaling: exact
        const unsigned long long dg = to_digits<NS>(w);
        const int off = (k >> 4) * (TN / 8) * 128 + (jj >> 3) * 128 + (jj & 7) * 16 + (k & 15);
#pragma unroll
        for (int s = 0; s < NS; ++s) base[(size_t)(NS - 1 - s) * TN * kI8H + off] = (unsigned char)(dg >> (8 * s));
    }
}

// ---- main kernel ------------------------------------------------------------------------------------------------------
// Persistent: one CTA per SM loops over 128-row blocks; per tile of TN output columns the NS(NS+1)/2 slice pairs are
// issued W2-slice-major (a slice is released as soon as its NS-q MMAs are done), then the epilogue drains the NS level
// accumulators.  MMA and tcgen05.ld are deliberately NOT overlapped: measured on B200 they serialise on TMEM anyway
// (free-running MMA 0.53 ms + free-running LDTM 0.40 ms = 0.98 ms together, profiles/), so a double-buffered or
// level-major accumulator scheme only adds hand-shakes.  What does overlap with the next tile's MMAs is the fp64
// scale/bias and the global stores, which run after TMEM has been handed back.
#ifndef JF_I8_DBG
#define JF_I8_DBG 0
#endif
template <typename T, int NS, int TN, int KR>
__global__ void __launch_bounds__(kI8Threads, 1) mlp2_i8_kernel(const __grid_constant__ MlpArgs<T> m,
                                                                 const unsigned char* __restrict__ wsB, int n_slots) {
    // timing experiments only (variant builds with -DJF_I8_DBG=<bits>, never the product): 1 no MMAs, 2 loader free-runs,
    // 4 no global stores, 8 no TMEM reads
    constexpr int dbg = JF_I8_DBG;
    using Cfg = I8Cfg<NS, TN>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const int Kin = m.dims[0], N = m.dims[2];
    const int n_tiles = (N + TN - 1) / TN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_blocks = (m.B + kI8Rows - 1) / kI8Rows;
    const int my_blocks = (int)((n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t offB = Cfg::offB, offBar = offB + n_slots * Cfg::kSliceBytesB;
    const uint32_t bar0 = sbase + offBar;
    // barriers (8 B each): full[16] | empty[16] | lvl_full[8] | lvl_empty[8] | a_free ; tmem pointer at +448
    auto bar_full = [&](int sl) { return bar0 + 8 * sl; };
    auto bar_empty = [&](int sl) { return bar0 + 8 * (16 + sl); };
    auto bar_lvl_full = [&](int l) { return bar0 + 8 * (32 + l); };
    auto bar_lvl_empty = [&](int l) { return bar0 + 8 * (40 + l); };
    const uint32_t bar_a_free = bar0 + 8 * 48;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + offBar + 448);
    double* sE = reinterpret_cast<double*>(smem + offBar + 512);   // [2][2][TN]: scale, b2 of the current / next tile
    T* sW1 = reinterpret_cast<T*>(sE + 4 * TN);                    // [Kin][128]
    T* sB1 = sW1 + (size_t)Kin * kI8H;
    T* sIn = sB1 + kI8H;                                           // [128][Kin|1]
    const int ldin = Kin | 1;
    // W2 tiles are visited in a per-CTA rotated order so that the CTAs do not all hit the same L2 lines at once
    const int tile_rot = (int)(blockIdx.x % (unsigned)n_tiles);
    auto tile_of = [&](int t) { int x = t + tile_rot; return x >= n_tiles ? x - n_tiles : x; };
    const int loads_per_block = n_tiles * NS;
    const int64_t total_loads = (int64_t)my_blocks * loads_per_block;
    // producer state (warp 1, lane 0): running slice index, its ring slot / wrap count, its (tile, q) within a block
    int64_t next_load = 0;
    int ld_slot = 0, ld_wrap = 0, ld_t = 0, ld_q = 0;
    auto issue_load = [&]() {                                       // issue slice `next_load` and advance the state
        if (dbg & 2) mbar_arrive(bar_full(ld_slot));
        else {
            mbar_expect_tx(bar_full(ld_slot), Cfg::kSliceBytesB);
            bulk_g2s(sbase + offB + ld_slot * Cfg::kSliceBytesB, wsB + ((size_t)tile_of(ld_t) * NS + ld_q) * Cfg::kSliceBytesB,
                     Cfg::kSliceBytesB, bar_full(ld_slot));
        }
        ++next_load;
        if (++ld_slot == n_slots) { ld_slot = 0; ++ld_wrap; }
        if (++ld_q == NS) { ld_q = 0; if (++ld_t == n_tiles) ld_t = 0; }
    };

    // ---- one-time setup: barriers, TMEM, W1^T ---------------------------------------------------------------------------
    if (warp == 1 && lane == 0) {
        for (int sl = 0; sl < n_slots; ++sl) { mbar_init(bar_full(sl), 1); mbar_init(bar_empty(sl), 1); }
        for (int l = 0; l < NS; ++l) { mbar_init(bar_lvl_full(l), 1); mbar_init(bar_lvl_empty(l), kI8EpiWarps); }
        mbar_init(bar_a_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(bar0 + 448), "n"(Cfg::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < Kin * kI8H; e += kI8Threads) {
        const int i = e / kI8H, u = e - i * kI8H;
        sW1[e] = m.wt[0][(size_t)u * Kin + i];
    }
    if (tid < kI8H) sB1[tid] = m.bias[0][tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 1 && lane == 0) {                                    // the first ring of W2 slices lands during the prologue
        while (next_load < n_slots - 1 && next_load < total_loads) issue_load();
    }
    // MMA issuer state (warp 0, elected lane): ring slot / wrap count of slice q = 0 of the current tile, tile parity
    int mm_slot0 = 0, mm_wrap0 = 0;
    uint32_t tile_par = 0;                                           // parity of the running tile index (all roles)
    const double* scl_g = reinterpret_cast<const double*>(wsB + (size_t)n_tiles * NS * TN * kI8H);
    const T* __restrict__ b2_g = m.bias[1];
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(kI8Rows >> 4) << 24);

#pragma unroll 1
    for (int jb = 0; jb < my_blocks; ++jb) {
        const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)jb * gridDim.x) * kI8Rows;
        // ---- gather the concatenated input rows of this block ----
        for (int e = tid; e < kI8Rows * Kin; e += kI8Threads) {
            const int r = e / Kin;
            int c = e - r * Kin;
            const int64_t row = row0 + r;
            T v = T(0);
            if (row < m.B) {
                int sg = 0;
                while (c >= m.seg_cols[sg]) { c -= m.seg_cols[sg]; ++sg; }
                v = m.seg_ptr[sg][row * m.seg_ld[sg] + c];
            }
            sIn[r * ldin + (e - r * Kin)] = v;
        }
        __syncthreads();
        if (jb > 0) mbar_wait(bar_a_free, (jb - 1) & 1);              // the MMAs of the previous block have read A
        // ---- prologue: layer 1 + tanh + digits -> A slices ----
        if (!(dbg & 8)) {
            const int r = (warp & 3) * 32 + lane, quarter = warp >> 2;    // thread = (row, quarter of the hidden units)
            T in[KR];                                                      // fp64 path only (Kin <= KR <= 16)
            if (sizeof(T) == 8) {
#pragma unroll
                for (int i = 0; i < KR; ++i) in[i] = (i < Kin) ? sIn[r * ldin + i] : T(0);
            }
            const uint32_t a_row = sbase + Cfg::offA + (r >> 3) * 128 + (r & 7) * 16;
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {                              // 16 hidden units per chunk
                const int chunk = quarter * 2 + ch;
                uint32_t lo[4][4], hi[4][4];                               // [group of 4 units][digit within word]
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    unsigned long long dg[4];
                    if (sizeof(T) == 8) {
#pragma unroll
                        for (int uu = 0; uu < 4; ++uu) {
                            const int u = chunk * 16 + gq * 4 + uu;
                            T z = sB1[u];
#pragma unroll
                            for (int i = 0; i < KR; ++i)
                                if (i < Kin) z = fma(in[i], sW1[i * kI8H + u], z);
                            dg[uu] = to_digits<NS>(tanh_abs(z));
                        }
                    } else {
                        // fp32: many inputs (conditional input + embeddings): stream them from shared memory, 4 units at once
                        const int u0 = chunk * 16 + gq * 4;
                        T z0 = sB1[u0], z1 = sB1[u0 + 1], z2 = sB1[u0 + 2], z3 = sB1[u0 + 3];
                        const T* xin = sIn + r * ldin;
#pragma unroll 4
                        for (int i = 0; i < Kin; ++i) {
                            const T xi = xin[i];
                            const T* wrow = sW1 + i * kI8H + u0;
                            z0 = fma(xi, wrow[0], z0); z1 = fma(xi, wrow[1], z1);
                            z2 = fma(xi, wrow[2], z2); z3 = fma(xi, wrow[3], z3);
                        }
                        dg[0] = to_digits<NS>(tanh_abs(z0)); dg[1] = to_digits<NS>(tanh_abs(z1));
                        dg[2] = to_digits<NS>(tanh_abs(z2)); dg[3] = to_digits<NS>(tanh_abs(z3));
                    }
                    // 4x4 byte transposes: word s of the result holds digit s of units 0..3 (byte b = unit b)
                    {
                        const uint32_t a0 = (uint32_t)dg[0], a1 = (uint32_t)dg[1], a2 = (uint32_t)dg[2], a3 = (uint32_t)dg[3];
                        const uint32_t x01 = __byte_perm(a0, a1, 0x5140), y01 = __byte_perm(a0, a1, 0x7362);
                        const uint32_t x23 = __byte_perm(a2, a3, 0x5140), y23 = __byte_perm(a2, a3, 0x7362);
                        lo[gq][0] = __byte_perm(x01, x23, 0x5410); lo[gq][1] = __byte_perm(x01, x23, 0x7632);
                        lo[gq][2] = __byte_perm(y01, y23, 0x5410); lo[gq][3] = __byte_perm(y01, y23, 0x7632);
                    }
                    if (NS > 4) {
                        const uint32_t a0 = (uint32_t)(dg[0] >> 32), a1 = (uint32_t)(dg[1] >> 32),
                                       a2 = (uint32_t)(dg[2] >> 32), a3 = (uint32_t)(dg[3] >> 32);
                        const uint32_t x01 = __byte_perm(a0, a1, 0x5140), y01 = __byte_perm(a0, a1, 0x7362);
                        const uint32_t x23 = __byte_perm(a2, a3, 0x5140), y23 = __byte_perm(a2, a3, 0x7362);
                        hi[gq][0] = __byte_perm(x01, x23, 0x5410); hi[gq][1] = __byte_perm(x01, x23, 0x7632);
                        hi[gq][2] = __byte_perm(y01, y23, 0x5410); hi[gq][3] = __byte_perm(y01, y23, 0x7632);
                    }
                }
#pragma unroll
                for (int sd = 0; sd < NS; ++sd) {                         // digit sd -> slice p = NS-1-sd
                    const uint32_t addr = a_row + (NS - 1 - sd) * Cfg::kSliceBytesA + chunk * Cfg::kLboA;
                    uint32_t w0, w1, w2, w3;
                    if (sd < 4) { w0 = lo[0][sd]; w1 = lo[1][sd]; w2 = lo[2][sd]; w3 = lo[3][sd]; }
                    else { w0 = hi[0][sd & 3]; w1 = hi[1][sd & 3]; w2 = hi[2][sd & 3]; w3 = hi[3][sd & 3]; }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncthreads();

        // ---- main phase of this block: warp-specialised ----
        if (warp == 1) {
            if (lane == 0) {
                // keep the ring full; run ahead into the next block's first slices while the last tile computes
                int64_t lim = (int64_t)(jb + 1) * loads_per_block + (n_slots - NS - 1);
                if (lim > total_loads) lim = total_loads;
                while (next_load < lim) {
                    mbar_wait(bar_empty(ld_slot), (uint32_t)((ld_wrap & 1) ^ 1));
                    issue_load();
                }
            }
        } else if (warp == 0) {
            if (elect_one()) {
                const uint32_t desc_hi = (Cfg::kSbo >> 4) | (1u << 14);                     // SBO, descriptor version 1
                const uint32_t a_lo0 = (((sbase + Cfg::offA) & 0x3FFFF) >> 4) | ((uint32_t)(Cfg::kLboA >> 4) << 16);
                const uint32_t b_lo0 = (((sbase + offB) & 0x3FFFF) >> 4) | ((uint32_t)(Cfg::kLboB >> 4) << 16);
                uint32_t par = tile_par;
                for (int t = 0; t < n_tiles; ++t, par ^= 1u) {
                    mbar_wait(bar_lvl_empty(0), par ^ 1u);           // the epilogue has drained the accumulators of the previous tile
                    tc_fence_after();
#pragma unroll
                    for (int q = 0; q < NS; ++q) {
                        int sl = mm_slot0 + q, wr = mm_wrap0;
                        if (sl >= n_slots) { sl -= n_slots; ++wr; }
                        mbar_wait(bar_full(sl), (uint32_t)(wr & 1));                  // W2 slice q of this tile landed
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + sl * (Cfg::kSliceBytesB >> 4);
                        if (!(dbg & 1)) {
#pragma unroll
                            for (int p = 0; p + q < NS; ++p)
                                tc_mma_i8_x4(tmem + (p + q) * TN, a_lo0 + p * (Cfg::kSliceBytesA >> 4), b_lo, desc_hi, desc_hi, idesc,
                                             q > 0 ? 1u : 0u, (2 * Cfg::kLboA) >> 4, (2 * Cfg::kLboB) >> 4);
                        }
                        tc_commit(bar_empty(sl));                    // the slot is free once these MMAs have read it
                    }
                    tc_commit(bar_lvl_full(0));                      // all levels of this tile are complete
                    mm_slot0 += NS;
                    if (mm_slot0 >= n_slots) { mm_slot0 -= n_slots; ++mm_wrap0; }
                }
                tc_commit(bar_a_free);
            }
        } else if (warp >= 4) {
            const int lq = warp & 3, grp = (warp - 4) >> 2;   // TMEM lane quarter (a warp may touch lanes 32*(w%4)..+31), column group
            const int r = lq * 32 + lane;                     // TMEM lane = row
            const int et = tid - 128;                         // 0..383 within the epilogue group
            const int64_t row = row0 + r;
            constexpr int kChunks = TN / 8, kGroups = kI8EpiWarps / 4, kMine = (kChunks + kGroups - 1) / kGroups;
            auto fetch_consts = [&](int t) {                  // scale and b2 of tile t -> sE[t & 1] with cp.async (non-blocking)
                if (et < 2 * TN) {
                    const int which = et / TN, jj = et - which * TN;
                    const int n = tile_of(t) * TN + jj;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sE + ((t & 1) * 2 + which) * TN + jj);
                    if (which == 0) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(scl_g + n) : "memory");
                    } else {        // b2 has the caller's element size: 8 or 4 bytes into the 8-byte cell
                        const T* src = b2_g + (n < N ? n : N - 1);
                        if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
                        else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            fetch_consts(0);
            uint32_t par = tile_par;
            for (int t = 0; t < n_tiles; ++t, par ^= 1u) {
                if (t + 1 < n_tiles) fetch_consts(t + 1); else asm volatile("cp.async.commit_group;" ::: "memory");
                T acc_d[kMine][8];
                const uint32_t tbase = tmem + ((uint32_t)(lq * 32) << 16);
                mbar_wait(bar_lvl_full(0), par);
                tc_fence_after();
#pragma unroll
                for (int i = 0; i < kMine; ++i) {
                    const int c = grp + i * kGroups;
                    if (c < kChunks && !(dbg & 4)) {
                        int v[NS][8];
#pragma unroll
                        for (int l = 0; l < NS; ++l) tmem_ld8(tbase + l * TN + c * 8, v[l]);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (sizeof(T) == 8) {
                                // Horner over the levels in fp64; int32 -> double without I2F: the bit pattern of 2^52 + 2^31 + v
                                // minus that bias is exact.  7 roundings at 2^-53 relative: far below the 2^-8NS truncation.
                                double sacc = __hiloint2double(0x43300000, v[NS - 1][j] ^ 0x80000000) - 4503601774854144.0;
#pragma unroll
                                for (int l = NS - 2; l >= 0; --l)
                                    sacc = fma(sacc, 0.00390625, __hiloint2double(0x43300000, v[l][j] ^ 0x80000000) - 4503601774854144.0);
                                acc_d[i][j] = (T)sacc;
                            } else {
                                // fp32: the level sums (< 2^24) are exact floats; Horner in fp32 (NS roundings at 2^-24)
                                float sacc = (float)v[NS - 1][j];
#pragma unroll
                                for (int l = NS - 2; l >= 0; --l) sacc = fmaf(sacc, 0.00390625f, (float)v[l][j]);
                                acc_d[i][j] = (T)sacc;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_lvl_empty(0));        // TMEM is free: the next tile's MMAs may start
                asm volatile("cp.async.wait_group 1;" ::: "memory");        // constants of tile t (issued one tile ago)
                asm volatile("bar.sync 1, 384;" ::: "memory");              // ... written by other epilogue threads
                const double* sc = sE + ((t & 1) * 2) * TN;
                const int n0 = tile_of(t) * TN;
#pragma unroll
                for (int i = 0; i < kMine; ++i) {
                    const int c = grp + i * kGroups;
                    if (c < kChunks && !(dbg & 4)) {
                        // one 64-bit address per chunk of 8 columns, then a walking pointer: the first version computed
                        // n * so_p + row * so_r per element -- 36 % of the kernel's executed instructions (ncu source page)
                        const int nb = n0 + c * 8;
                        T* op = m.out + (int64_t)nb * m.so_p + row * m.so_r;
                        const int nvalid = (row < m.B) ? (N - nb) : 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            T o;
                            if (sizeof(T) == 8) o = (T)fma((double)acc_d[i][j], sc[c * 8 + j], sc[TN + c * 8 + j]);
                            else o = (T)fmaf((float)acc_d[i][j], reinterpret_cast<const float*>(sc + c * 8 + j)[0],
                                             reinterpret_cast<const float*>(sc + TN + c * 8 + j)[0]);
                            if (j < nvalid) *op = o;
                            op += m.so_p;
                        }
                    }
                }
                asm volatile("bar.sync 1, 384;" ::: "memory");              // sE[t & 1] may be refilled (for tile t + 2)
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        tile_par ^= (uint32_t)(n_tiles & 1);
    }
    // ---- teardown -------------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(Cfg::kTmemCols) : "memory");
    }
}

}  // namespace jf
