// Fused parameter generator + "g" chain: one kernel computes the per-row flow parameters of a conditional
// Euclidean sub-pdf on the tensor cores AND consumes them in the layer chain, so the [P, rows] parameter block (4.4 KB
// per row for the README flow) never exists in HBM.
//
//   reference hand-off: main/default.py:956 (MLP call) -> :998-1029 (layer loop, `extra_inputs` slices), sampling
//   :1438 -> :1482-1506; layer math gaussianization_flow.py:389-1114 (same arithmetic as csrc/gf.cuh, register-resident).
//
// Structure (one persistent CTA per SM, 128 rows per block, 20 warps in 5 warpgroups):
//   warps 0-15  workers.  Prologue: layer 1 + tanh + int8 digits of the hidden
//               activations, written with tcgen05.st into TENSOR MEMORY: the A operand of the MMAs lives in TMEM (32
//               columns per int8 slice), which takes the A read off the shared-memory port (an SS-mode M = 128 MMA costs
//               33 cycles whatever N is; A-in-TMEM N = 48 costs 24) and frees 96 KB of shared memory for the staging
//               below.  Then worker (row, j) owns DIMENSION j of its row: per layer it takes ITS parameter columns of
//               each tile from the staging buffer into registers and consumes them:
//                 log_pdf   STREAMING: a tile carries (mean, log-width, log-norm) TRIPLES, each kernel is regulated and
//                           added to the mixture sums as it arrives, nothing but the sums stays live (tile 0 opens with the
//                           Householder components and the offset, so the rotated coordinate is known before the first
//                           kernel); an online rescaling exponent keeps the sums exact however far out x is;
//                 sampling  the K regulated kernels of the dimension go to the worker's shared-memory slots (the 96 KB that
//                           the A operand left free), where the root finder of csrc/gf.cuh reads them.
//               The four workers of a row meet once per layer through a 6-field shared-memory exchange for the rotation.
//   warps 16-19 the tensor warpgroup.  One elected thread streams the pre-sliced W2 tiles (L2 resident) with
//               cp.async.bulk into a two-tile ring and issues the tcgen05.mma (kind::i8, M = 128, K = 32, A from TMEM) of a
//               tile; then all four warps -- one per TMEM lane quarter -- drain the level accumulators: tcgen05.ld, exact
//               64-bit INTEGER combine of the levels (ALU pipe), one conversion, scale and bias, and the fp64 parameters
//               go to the staging buffer in shared memory.  MMA and TMEM reads serialise on tensor memory anyway (measured,
//               profiles/), so one warpgroup doing both loses nothing, and it runs a tile or two AHEAD of the workers:
//               the FP64 pipe (the bound of this kernel) never waits for a TMEM read (64 B/clk per SM: 1.2 us per tile).
// W2's rows are permuted by the prep kernel into CONSUMPTION order (per direction): tile (c, part) holds, for the layer
// consumed c-th, the SPT = TN/4 columns of `part` for each of the 4 dimensions.
#pragma once
#include <cstdio>
#include "mlp_i8.cuh"
#include "gf.cuh"
#include "gf_fused_launch.cuh"

namespace jf {

constexpr int kFuWorkers = 512;     // 16 warps
constexpr int kFuThreads = 640;     // + the tensor warpgroup
constexpr int kFuExFields = 6;      // exchange: x, 4 Householder components, log-derivative
constexpr int kFuExBufBytes = kFuMaxD * kFuExFields * kI8Rows * 8;
// the level accumulators of one output are combined as ONE 64-bit integer: sum_{l < 6} v_l 256^(5-l) < 2^62
// (|v_l| <= (l+1) 2^21); a 7th level is added in floating point
template <int NS> struct FuLv { static constexpr int kInt = NS < 6 ? NS : 6; };
// timing experiments only (tools/fused_variants.py builds variant libraries; results are garbage with any bit set):
// 1 no MMAs, 2 no layer arithmetic in the workers, 4 no layer-1 / tanh arithmetic, 8 no TMEM reads.  Product build: 0.
#ifndef JF_FU_DBG
#define JF_FU_DBG 0
#endif
// JF_FU_PROF=1 (variant builds only): lane 0 of every tensor warp of CTA 0 accumulates clock64() per phase and prints them
#ifndef JF_FU_PROF
#define JF_FU_PROF 0
#endif
#if JF_FU_PROF
#define FU_T(i) do { const long long now_ = clock64(); prof_[i] += now_ - last_; last_ = now_; } while (0)
#else
#define FU_T(i) do { } while (0)
#endif

// geometry of a (slices, tile width, direction) configuration
template <int NS, int TN, int DIR = JF_DIR_LOGPDF>
struct FuCfg {
    static constexpr int kSPT = TN / kFuMaxD;                 // parameter slots per (tile, dimension)
    static constexpr int kTPL = (36 + kSPT - 1) / kSPT;       // tiles per layer: 36 slots per (layer, dimension) for K = 10
    static constexpr int kAccCol = 0;                         // TMEM: NS level accumulators of TN columns ...
    static constexpr int kACol = NS * TN;                     // ... then the NS A slices of 32 columns (128 int8 per row)
    static_assert(kACol + NS * 32 <= 512, "tensor memory: 512 columns");
    static_assert(TN % 16 == 0 && TN % kFuMaxD == 0, "UMMA N for M = 128; equal share per dimension");
    static constexpr int kSliceBytesB = TN * kI8H;
    static constexpr int kLboB = (TN / 8) * 128, kSbo = 128;
    static constexpr int kRing = NS;                          // W2 slices of one tile (refilled slice by slice as the MMAs retire)
    static constexpr int kStageBytes = TN * kI8Rows * 8;      // fp64 parameters of one tile, [column][row]
    static constexpr int kConstBytes = 3 * TN * 16;           // (scale, b2) of the tile being drained, the next one, and one in between
                                                              // (a warp may still store the last columns of the previous tile)
    // sampling keeps the regulated kernels of every (row, dimension) in shared-memory slots for the root finder
    // ([field m / 1/w / n][k][worker]: 120 KB); to make room its exchange lives inside the slots (two barriers per
    // meeting) and W1 / the gathered inputs -- prologue only -- live inside staging buffer 0
    static constexpr bool kSample = DIR == JF_DIR_SAMPLE;
    static constexpr int kExBufs = kSample ? 0 : 2;          // sampling: the exchange lives in the slots
    static constexpr int kSlotBytes = kSample ? 3 * kFuK * kFuWorkers * 8 : 0;
    static constexpr int offB = 0;
    static constexpr int offStage = offB + kRing * kSliceBytesB;
    __host__ __device__ static constexpr int pro_bytes(int kin) { return (kin + 1) * kI8H * 8 + kI8Rows * (kin | 1) * 8; }   // W1^T, b1, inputs
    __host__ __device__ static constexpr int off_bar(int n_stages) { return offStage + n_stages * kStageBytes; }
    __host__ __device__ static constexpr int off_ex(int n_stages) { return off_bar(n_stages) + 512 + kConstBytes; }
    __host__ __device__ static constexpr int off_slots(int n_stages) { return off_ex(n_stages) + kExBufs * kFuExBufBytes; }
    __host__ __device__ static constexpr int off_pro(int n_stages) { return kSample ? offStage : off_slots(n_stages) + kSlotBytes; }
    __host__ __device__ static constexpr int smem_bytes(int kin, int n_stages) {
        return kSample ? off_slots(n_stages) + kSlotBytes : off_pro(n_stages) + pro_bytes(kin);
    }
    // (sampling: W1 and the inputs must fit into one staging buffer)
    __host__ __device__ static constexpr bool fits(int kin, int n_stages, int smem_max) {
        return smem_bytes(kin, n_stages) <= smem_max && (!kSample || pro_bytes(kin) <= kStageBytes);
    }
};

template <int NS, int TN>
__host__ __device__ inline int64_t fu_prep_bytes(int n_layers) {
    const int64_t n_tiles = (int64_t)FuCfg<NS, TN>::kTPL * n_layers;
    return n_tiles * NS * TN * kI8H + n_tiles * TN * 16;
}

// source row of W2 / b2 (index into the sub-pdf's raw parameter vector) of slot s of (layer, dimension cg); -1: zero column.
//   sampling  w[0..9], n[0..9], mean[0..9], v_0[j] .. v_3[j], offset_j, padding   (widths and norms first: their regulators
//             run while the tensor pipe produces the rest; everything the root finder needs sits in the first four
//             tiles, so the solve starts while the tile with the reflection components and the offset -- only needed
//             AFTER the solve -- is still being produced)
//   log_pdf   v_0[j] .. v_3[j], offset_j, (mean, log_w, log_n)[0], [1], one pad, then the triples 2..9 (12 slots per tile:
//             no triple straddles a tile)
__host__ __device__ inline int fu_source_param(const FuLayerC& c, int d, int direction, int cg, int s) {
    if (cg >= d) return -1;
    const int K = kFuK;
    const int off_hh = c.raw_off + (c.has_offset ? d : 0);
    const int off_m = off_hh + c.hh_iter * d, off_w = off_m + K * d, off_n = off_w + K * d;
    if (direction == JF_DIR_LOGPDF) {
        if (s < 4) return s < c.hh_iter ? off_hh + s * d + cg : -1;
        if (s == 4) return c.has_offset ? c.raw_off + cg : -1;
        if (s == 11 || s >= 36) return -1;
        const int t = s < 11 ? s - 5 : s - 6;          // index into the stream of triples
        const int k = t / 3, f = t - 3 * k;
        return (f == 0 ? off_m : (f == 1 ? off_w : off_n)) + k * d + cg;
    }
    if (s < 10) return off_w + s * d + cg;
    if (s < 20) return off_n + (s - 10) * d + cg;
    if (s < 30) return off_m + (s - 20) * d + cg;
    if (s < 34) return (s - 30) < c.hh_iter ? off_hh + (s - 30) * d + cg : -1;
    if (s == 34) return c.has_offset ? c.raw_off + cg : -1;
    return -1;
}

// W2 [P,128] fp64 -> int8 slices of the fused tiles (UMMA K-major no-swizzle layout) + (scale, b2) per fused column
template <int NS, int TN>
__global__ void __launch_bounds__(128) fu_prep_kernel(const __grid_constant__ FuArgs a, const double* __restrict__ W2,
                                                      const double* __restrict__ b2, int direction, unsigned char* ws) {
    using G = FuCfg<NS, TN>;
    const int tile = blockIdx.x, n_tiles = gridDim.x;
    const int c = tile / G::kTPL, part = tile - G::kTPL * c;
    const int l = direction == JF_DIR_LOGPDF ? a.n_layers - 1 - c : c;
    double2* cst = reinterpret_cast<double2*>(ws + (size_t)n_tiles * NS * TN * kI8H);
    __shared__ double s_inv[TN];
    __shared__ int s_src[TN];
    for (int jj = threadIdx.x; jj < TN; jj += blockDim.x) {
        const int cg = jj / G::kSPT, s = part * G::kSPT + (jj - cg * G::kSPT);
        const int src = fu_source_param(a.layers[l], a.d, direction, cg, s);
        double mx = 0.0;
        if (src >= 0)
            for (int k = 0; k < kI8H; ++k) mx = fmax(mx, fabs(W2[(size_t)src * kI8H + k]));
        int e = 0;
        if (mx > 0.0) { frexp(mx, &e); }
        s_inv[jj] = ldexp(1.0, -e);
        s_src[jj] = src;
        cst[tile * TN + jj] = src >= 0 ? make_double2(ldexp(1.0, e - 12 - 8 * (FuLv<NS>::kInt - 1)), b2[src]) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    unsigned char* base = ws + (size_t)tile * NS * TN * kI8H;
    for (int idx = threadIdx.x; idx < TN * kI8H; idx += blockDim.x) {
        const int jj = idx / kI8H, k = idx - jj * kI8H;
        const int src = s_src[jj];
        const double w = src >= 0 ? W2[(size_t)src * kI8H + k] * s_inv[jj] : 0.0;
        const unsigned long long dg = to_digits<NS>(w);
        const int off = (k >> 4) * (TN / 8) * 128 + (jj >> 3) * 128 + (jj & 7) * 16 + (k & 15);
#pragma unroll
        for (int s = 0; s < NS; ++s) base[(size_t)(NS - 1 - s) * TN * kI8H + off] = (unsigned char)(dg >> (8 * s));
    }
}

JF_DEVINL void tmem_ld4(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
JF_DEVINL void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// sum_{l < LI} v_l 256^(LI-1-l) as one 64-bit integer: the three lowest levels above the last go in with one
// mad.wide.s32 each (32 x 32 + 64 -> 64), the rest are additions to the high word
JF_DEVINL long long mad_wide(int a, int b, long long c) {
    long long d;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
template <int LI, int NSL, int CW>
JF_DEVINL long long fu_combine(const int (&lv)[NSL][CW], int j) {
    long long acc = (long long)lv[LI - 1][j];
    if (LI >= 2) acc = mad_wide(lv[LI - 2][j], 1 << 8, acc);
    if (LI >= 3) acc = mad_wide(lv[LI - 3][j], 1 << 16, acc);
    if (LI >= 4) acc = mad_wide(lv[LI - 4][j], 1 << 24, acc);
    if (LI >= 5) {
        int hi = (int)(acc >> 32) + lv[LI - 5][j];
        if (LI >= 6) hi += lv[LI - 6][j] << 8;
        acc = (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned long long)(unsigned)acc);
    }
    return acc;
}
JF_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T, int8 x int8 -> int32, M = 128: the four K = 32 steps of one (slice p, slice q) pair; A
// advances 8 columns (32 int8 per row) per step, the B descriptor by adding to its low word
JF_DEVINL void tc_mma_i8_ts_x4(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi_b, uint32_t idesc,
                               uint32_t acc_first, uint32_t b_step) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 db;\n.reg .b32 bl, al;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %4, p;\n"
        "setp.ne.b32 p, 1, 0;\n"
        "add.u32 al, %1, 8;\nadd.u32 bl, %2, %6;\nmov.b64 db, {bl, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [al], db, %4, p;\n"
        "add.u32 al, al, 8;\nadd.u32 bl, bl, %6;\nmov.b64 db, {bl, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [al], db, %4, p;\n"
        "add.u32 al, al, 8;\nadd.u32 bl, bl, %6;\nmov.b64 db, {bl, %3};\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [al], db, %4, p;\n}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(desc_hi_b), "r"(idesc), "r"(acc_first), "r"(b_step) : "memory");
}

// ---- log_pdf: streaming mixture sums (same quantities as mix_eval of csrc/gf.cuh, accumulated kernel by kernel) ------
// The rescaling exponent D plays the role of mix_eval's delta, but is found ONLINE: it is 0 unless the kernels seen so
// far are all more than 512 widths away from x, and whenever a nearer kernel arrives the scaled sums are brought to the
// new exponent (a rare, divergent branch).  With D = 0 -- every row that is not an extreme outlier -- the arithmetic is
// that of mix_eval with delta = 0; the norms are used unnormalised and the sums divided by their total at the end.
struct FuSums {
    double D, E;                                  // rescaling exponent of the "small" sums and of Sp; E = exp(-D)
    double big_p, small_p, big_n, small_n, Sp, ex, qc, nsum;
    int n_pos;                                    // kernels with a >= 0
};

// N kernels at once, branch-free in the common case so that the N (x 2: width and norm regulators) exp / reciprocal
// chains interleave: with four warps per scheduler the FP64 pipe is latency bound on a single chain (ncu: "wait").
// t: N triples (mean, raw log-width, raw log-norm).
template <int N, bool FIRST>
JF_DEVINL void fu_stream(FuSums& A, const FuLayerC& lc, double x, const double* t) {
    double iw[N], n[N], a[N], sa[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        iw[i] = regulate_inv_width(t[3 * i + 1], lc.w_min, lc.inv_w_max);
        n[i] = regulate_norm(t[3 * i + 2], lc.n_min, lc.n_max);
    }
    double smin = Num<double>::big, amin = 0.0, nsum = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        a[i] = (x - t[3 * i]) * iw[i];
        sa[i] = fabs(a[i]);
        smin = tmin(smin, sa[i]);
        amin = tmin(amin, a[i]);
        nsum += n[i];
    }
    if (FIRST) {
        A.D = smin > 512.0 ? smin : 0.0;
        A.E = 1.0;
        if (A.D > 0.0) A.E = exp_neg(-A.D);
        A.big_p = A.small_p = A.big_n = A.small_n = A.Sp = A.ex = A.qc = 0.0;
        A.nsum = nsum;
        A.n_pos = 0;
    } else {
        A.nsum += nsum;
        if (A.D > 0.0 && smin < A.D) {              // a nearer kernel than any before (extreme outliers only)
            const double nd = smin > 512.0 ? smin : 0.0;
            const double f = exp_neg(nd - A.D);
            A.small_p *= f; A.small_n *= f; A.Sp *= f; A.qc *= f;
            A.D = nd;
            A.E = nd > 0.0 ? exp_neg(-nd) : 1.0;
        }
    }
    double u[N], rx[N];
#pragma unroll
    for (int i = 0; i < N; ++i) u[i] = exp_neg(A.D - sa[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) rx[i] = rcp_1to2(fma(u[i], A.E, 1.0));
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double nr = n[i] * rx[i], nur = nr * u[i];
        const bool pos = a[i] >= 0.0;
        A.big_p += pos ? nr : 0.0;
        A.small_p += pos ? nur : 0.0;
        A.big_n += pos ? 0.0 : nr;
        A.small_n += pos ? 0.0 : nur;
        A.n_pos += pos ? 1 : 0;
        A.Sp = fma(nur * iw[i], rx[i], A.Sp);
    }
    if (amin < -20.0) {                             // softplus-threshold quirk of the reference, see mix_eval (rare)
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (a[i] < -20.0) {
                const double nq = n[i] * (u[i] * A.E) * rx[i];
                A.ex += nq;
                A.qc = fma(nq, u[i], A.qc);
                A.Sp = fma(nq * u[i] * iw[i], 1.0 + rx[i], A.Sp);
            }
    }
}

JF_DEVINL MixVal<double> fu_stream_finish(FuSums& A) {
    const bool all_neg = A.n_pos == 0, all_pos = A.n_pos == kFuK;
    if (A.D > 0.0 && !(all_neg || all_pos)) { A.small_n *= A.E; A.qc *= A.E; A.small_p *= A.E; }
    const double inv = rcp_pos_(A.nsum);
    MixVal<double> v;
    v.Sc = (A.big_p + A.small_n + A.qc) * inv;
    v.Ss = (A.small_p + A.big_n) * inv;
    v.Sp = A.Sp * inv;
    v.ex = A.ex * inv;
    v.Sd = 0.0;
    v.E = A.E;
    v.dc = all_neg ? A.D : 0.0;
    v.ds = all_pos ? A.D : 0.0;
    v.dp = A.D;
    return v;
}

// ---- main kernel ------------------------------------------------------------------------------------------------------
template <int NS, int TN, int DIR, int KR>
__global__ void __launch_bounds__(kFuThreads, 1) gf_fused_kernel(const __grid_constant__ FuArgs a, int n_stages) {
    using G = FuCfg<NS, TN, DIR>;
    constexpr int SPT = G::kSPT, TPL = G::kTPL;
    extern __shared__ __align__(1024) unsigned char smem[];
    const MlpArgs<double>& m = a.m;
    const int Kin = m.dims[0];
    const int L = a.n_layers, n_tiles = TPL * L, d = a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_blocks = (m.B + kI8Rows - 1) / kI8Rows;
    const int my_blocks = (int)((n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t offBar = G::off_bar(n_stages);
    const uint32_t bar0 = sbase + offBar;
    // barriers (8 B each): slice_full[8] | slice_empty[8] | mma_done | a_full | stage_full[4] | stage_empty[4] ; tmem pointer at +448
    auto bar_slice = [&](int sl) { return bar0 + 8 * sl; };
    auto bar_slice_empty = [&](int sl) { return bar0 + 8 * (8 + sl); };
    const uint32_t bar_mma_done = bar0 + 8 * 16, bar_a_full = bar0 + 8 * 17;
    auto bar_stage_full = [&](int s) { return bar0 + 8 * (18 + s); };
    auto bar_stage_empty = [&](int s) { return bar0 + 8 * (22 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + offBar + 448);
    double* sStage = reinterpret_cast<double*>(smem + G::offStage);        // [n_stages][TN][128]
    double2* sConst = reinterpret_cast<double2*>(smem + offBar + 512);     // [3][TN]
    double* sEx = reinterpret_cast<double*>(smem + G::off_ex(n_stages));   // [kExBufs][4][6][128]
    double* sSlots = reinterpret_cast<double*>(smem + G::off_slots(n_stages));   // sampling: [3][K][512]
    double* sW1 = reinterpret_cast<double*>(smem + G::off_pro(n_stages));  // [Kin][128]  (sampling: inside staging buffer 0)
    double* sB1 = sW1 + (size_t)Kin * kI8H;
    double* sIn = sB1 + kI8H;                                              // [128][Kin|1]
    const int ldin = Kin | 1;

    if (warp == 17 && lane == 0) {
        for (int sl = 0; sl < G::kRing; ++sl) { mbar_init(bar_slice(sl), 1); mbar_init(bar_slice_empty(sl), 1); }
        mbar_init(bar_mma_done, 1);
        mbar_init(bar_a_full, 16);
        for (int s = 0; s < n_stages; ++s) { mbar_init(bar_stage_full(s), 4); mbar_init(bar_stage_empty(s), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(bar0 + 448), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (!G::kSample) {
        for (int e = tid; e < Kin * kI8H; e += kFuThreads) {
            const int i = e / kI8H, u = e - i * kI8H;
            sW1[e] = m.wt[0][(size_t)u * Kin + i];
        }
        if (tid < kI8H) sB1[tid] = m.bias[0][tid];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(kI8Rows >> 4) << 24);

    if (warp >= 16) {
        // ---- tensor warpgroup: W2 ring, MMA issue, TMEM drain -> staging ----
        const int dq = warp - 16;
        const int r = dq * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(dq * 32) << 16);
        const int64_t total_tiles = (int64_t)my_blocks * n_tiles;
        auto issue_load = [&](int t, int q) {                 // slice q of tile t into ring slot q
            mbar_expect_tx(bar_slice(q), G::kSliceBytesB);
            bulk_g2s(sbase + G::offB + q * G::kSliceBytesB, a.wsB + ((size_t)t * NS + q) * G::kSliceBytesB, G::kSliceBytesB,
                     bar_slice(q));
        };
        if (warp == 16 && total_tiles > 0) {
            if (elect_one()) {
#pragma unroll
                for (int q = 0; q < NS; ++q) issue_load(0, q);
            }
        }
        const int tw = tid - 16 * 32;                                        // 0..127 inside the warpgroup
        if (tw < TN && total_tiles > 0) sConst[tw] = a.consts[tw];           // constants of tile 0
        bar_sync_named(7, 128);
#if JF_FU_PROF
        long long prof_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, last_ = clock64();
#endif
        int t = 0, jb = 0, stage = 0, cb = 0;                                // cb: constants buffer of the current tile (gt mod 3)
        uint32_t mma_par = 0, stage_wrap = 0;
#pragma unroll 1
        for (int64_t gt = 0; gt < total_tiles; ++gt) {
            // one elected thread of warp 16 (a warp-uniform branch + elect.sync: UTCIMMA takes uniform-register operands, and
            // under a divergent `lane == 0` the compiler wraps every single MMA in an election loop -- 51 cycles per MMA)
            FU_T(9);
            if (warp == 16 && elect_one()) {
                if (t == 0) mbar_wait(bar_a_full, (uint32_t)(jb & 1));        // the workers have written this block's A slices
                tc_fence_after();
                FU_T(0);
                const uint32_t desc_hi = (G::kSbo >> 4) | (1u << 14);
                const uint32_t b_lo0 = (((sbase + G::offB) & 0x3FFFF) >> 4) | ((uint32_t)(G::kLboB >> 4) << 16);
                const uint32_t ring_par = (uint32_t)(gt & 1);
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                    mbar_wait(bar_slice(q), ring_par);
                    tc_fence_after();
                    const uint32_t b_lo = b_lo0 + q * (G::kSliceBytesB >> 4);
#pragma unroll
                    for (int p = 0; p + q < NS && !(JF_FU_DBG & 1); ++p)
                        tc_mma_i8_ts_x4(tmem + G::kAccCol + (p + q) * TN, tmem + G::kACol + p * 32, b_lo, desc_hi, idesc,
                                        q > 0 ? 1u : 0u, (2 * G::kLboB) >> 4);
                    tc_commit(bar_slice_empty(q));               // slot q is free once these MMAs have read it
                }
                tc_commit(bar_mma_done);
                FU_T(1);
                // refill the ring with the next tile slice by slice as the MMAs retire (the last slot frees when the tile
                // is complete, which is when the drain below can start anyway)
                if (gt + 1 < total_tiles) {
                    const int tn = t + 1 == n_tiles ? 0 : t + 1;
#pragma unroll
                    for (int q = 0; q < NS; ++q) {
                        mbar_wait(bar_slice_empty(q), ring_par);
                        issue_load(tn, q);
                    }
                }
            }
            FU_T(2);
            __syncwarp();                                                    // (the issuer's 31 siblings wait here, not in a spin loop)
            FU_T(3);
            mbar_wait(bar_mma_done, mma_par); mma_par ^= 1u;
            tc_fence_after();
            FU_T(4);
            mbar_wait(bar_stage_empty(stage), stage_wrap ^ 1u);              // the workers have taken the previous content
            FU_T(5);
            {
                // software-pipelined drain, 4 columns per step in two register buffers: the tcgen05.ld of the next step are in
                // flight while the current step is combined and stored (TMEM reads are port bound: 64 B/clk per SM)
                constexpr int LI = FuLv<NS>::kInt;
                double* st = sStage + (size_t)stage * (TN * kI8Rows) + r;
                const double2* cst = sConst + cb * TN;
                int lvA[NS][4], lvB[NS][4];
                auto issue = [&](int (&lv)[NS][4], int c0) {
#pragma unroll
                    for (int l = 0; l < NS; ++l) {
                        if (JF_FU_DBG & 8) { lv[l][0] = lv[l][1] = lv[l][2] = lv[l][3] = l + c0; }
                        else tmem_ld4(tlane + G::kAccCol + l * TN + c0, lv[l]);
                    }
                };
                auto process = [&](const int (&lv)[NS][4], int c0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double sacc = __ll2double_rn(fu_combine<LI>(lv, j));
#pragma unroll
                        for (int l = LI; l < NS; ++l) sacc = fma((double)lv[l][j], 1.0 / (double)(1ull << (8 * (l - LI + 1))), sacc);
                        const double2 sb = cst[c0 + j];
                        st[(size_t)(c0 + j) * kI8Rows] = fma(sacc, sb.x, sb.y);
                    }
                };
                issue(lvA, 0);
                if (!(JF_FU_DBG & 8)) tmem_ld_wait();
#pragma unroll 1
                for (int c0 = 0; c0 < TN - 8; c0 += 8) {
                    issue(lvB, c0 + 4);
                    process(lvA, c0);
                    if (!(JF_FU_DBG & 8)) tmem_ld_wait();
                    issue(lvA, c0 + 8);
                    process(lvB, c0 + 4);
                    if (!(JF_FU_DBG & 8)) tmem_ld_wait();
                }
                issue(lvB, TN - 4);
                process(lvA, TN - 8);
                if (!(JF_FU_DBG & 8)) tmem_ld_wait();
                // constants of the next tile (read after the barrier below)
                if (tw < TN && gt + 1 < total_tiles)
                    sConst[(cb == 2 ? 0 : cb + 1) * TN + tw] = a.consts[(size_t)(t + 1 == n_tiles ? 0 : t + 1) * TN + tw];
                tc_fence_before();
                FU_T(6);
                bar_sync_named(7, 128);                                      // TMEM is free: the next tile's MMAs may start
                FU_T(7);
                process(lvB, TN - 4);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_stage_full(stage));
            FU_T(8);
            if (++stage == n_stages) { stage = 0; stage_wrap ^= 1u; }
            if (++t == n_tiles) { t = 0; ++jb; }
            cb = cb == 2 ? 0 : cb + 1;
        }
#if JF_FU_PROF
        if (blockIdx.x == 0 && lane == 0)
            printf("warp %d tiles %lld cycles/tile: a_full %lld | slices+mma issue %lld | refill(=mma run) %lld | syncwarp %lld | mma_done %lld | stage_empty %lld | drain %lld | bar7 %lld | tail %lld | loop %lld\n",
                   warp, total_tiles, prof_[0] / total_tiles, prof_[1] / total_tiles, prof_[2] / total_tiles, prof_[3] / total_tiles,
                   prof_[4] / total_tiles, prof_[5] / total_tiles, prof_[6] / total_tiles, prof_[7] / total_tiles, prof_[8] / total_tiles,
                   prof_[9] / total_tiles);
#endif
    } else {
        // ---- workers ----
        const int lq = warp & 3, cg = warp >> 2;
        const int r = lq * 32 + lane;
        const bool active = cg < d;
        const uint32_t tlane = tmem + ((uint32_t)(lq * 32) << 16);
        int wstage = 0;
        uint32_t wwrap = 0;
        int ex_buf = 0;
        int n_evals = 0, n_unconv = 0, n_bad = 0;
        // exchange of the four workers of a row: field f of dimension j at exr[buf_off + j * exJ + f * exF].  log_pdf: its own
        // double-buffered array; sampling: the workers' shared-memory slots (mean slots k = 0..5 of worker (row, j)), which
        // are dead between a worker's solve and the arrival of the next layer's means -- that is what frees 24 KB for a
        // second staging buffer
        double* exr = (G::kSample ? sSlots : sEx) + r;
        constexpr int exF = G::kSample ? kFuWorkers : kI8Rows, exJ = G::kSample ? kI8Rows : kFuExFields * kI8Rows;
#if JF_FU_PROF
        long long prof_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, last_ = clock64();
#endif
        // this worker's SPT values of the next tile: staging -> registers, then the stage goes back to the tensor warpgroup
        auto next_tile = [&](double* v) {
            FU_T(0);
            mbar_wait(bar_stage_full(wstage), wwrap);
            FU_T(1);
            const double* st = sStage + (size_t)wstage * (TN * kI8Rows) + (size_t)(cg * SPT) * kI8Rows + r;
#pragma unroll
            for (int i = 0; i < SPT; ++i) v[i] = st[(size_t)i * kI8Rows];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_stage_empty(wstage));
            if (++wstage == n_stages) { wstage = 0; wwrap ^= 1u; }
        };
#pragma unroll 1
        for (int jb = 0; jb < my_blocks; ++jb) {
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)jb * gridDim.x) * kI8Rows;
            const int64_t row = row0 + r;
            const bool live = row < m.B;
            if (G::kSample) {
                // W1 and the inputs live in staging buffer 0 (prologue only): every worker has taken the last tile of the
                // previous block, and the tensor warpgroup cannot write the buffer again before this block's A is complete
                bar_sync_named(1, kFuWorkers);
                for (int e = tid; e < Kin * kI8H; e += kFuWorkers) {
                    const int i = e / kI8H, u = e - i * kI8H;
                    sW1[e] = m.wt[0][(size_t)u * Kin + i];
                }
                if (tid < kI8H) sB1[tid] = m.bias[0][tid];
            }
            // ---- gather the generator's input rows ----
            for (int e = tid; e < kI8Rows * Kin; e += kFuWorkers) {
                const int rr = e / Kin;
                int c = e - rr * Kin;
                const int64_t grow = row0 + rr;
                double v = 0.0;
                if (grow < m.B) {
                    int sg = 0;
                    while (c >= m.seg_cols[sg]) { c -= m.seg_cols[sg]; ++sg; }
                    v = m.seg_ptr[sg][grow * m.seg_ld[sg] + c];
                }
                sIn[rr * ldin + (e - rr * Kin)] = v;
            }
            FU_T(2);
            bar_sync_named(1, kFuWorkers);
            FU_T(3);
            // ---- prologue: layer 1 + tanh + digits -> A slices in tensor memory (thread = (row, quarter of the hidden
            //      units); a 32-bit TMEM column holds 4 consecutive int8 of the row, slice p occupies columns 32 p .. 32 p + 31) ----
            {
                const int quarter = cg;
                double in[KR];
#pragma unroll
                for (int i = 0; i < KR; ++i) in[i] = (i < Kin) ? sIn[r * ldin + i] : 0.0;
#pragma unroll 1
                for (int ch = 0; ch < 2; ++ch) {
                    const int chunk = quarter * 2 + ch;
                    uint32_t lo[4][4], hi[4][4];
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        unsigned long long dg[4];
#pragma unroll
                        for (int uu = 0; uu < 4; ++uu) {
                            const int u = chunk * 16 + gq * 4 + uu;
                            double z = sB1[u];
#pragma unroll
                            for (int i = 0; i < KR; ++i)
                                if (i < Kin) z = fma(in[i], sW1[i * kI8H + u], z);
                            dg[uu] = (JF_FU_DBG & 4) ? (unsigned long long)(u + r) : to_digits<NS>(tanh_abs(z));
                        }
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
                    for (int sd = 0; sd < NS; ++sd) {                       // digit sd -> slice p = NS-1-sd
                        const uint32_t taddr = tlane + G::kACol + (NS - 1 - sd) * 32 + chunk * 4;
                        if (sd < 4) tmem_st4(taddr, lo[0][sd], lo[1][sd], lo[2][sd], lo[3][sd]);
                        else tmem_st4(taddr, hi[0][sd & 3], hi[1][sd & 3], hi[2][sd & 3], hi[3][sd & 3]);
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a_full);
            FU_T(4);

            // ---- the layer chain of this row, dimension cg ----
            double xj = (live && active) ? a.in[row * a.ld_in + cg] : 0.0;
            double logdet_acc = 0.0, zsq = 0.0;         // (worker 0 of the row only)
            if (cg == 0) {
                logdet_acc = (live && a.logdet_in) ? a.logdet_in[row] : 0.0;
                if (DIR == JF_DIR_SAMPLE && live)
                    for (int j = 0; j < d; ++j) { const double zj = a.in[row * a.ld_in + j]; zsq = fma(zj, zj, zsq); }
            }
            double logd_prev = 0.0;                     // log_pdf: log-derivative of the previous layer, not yet exchanged
#pragma unroll 1
            for (int c = 0; c < L; ++c) {
                const int l = DIR == JF_DIR_LOGPDF ? L - 1 - c : c;
                const FuLayerC& lc = a.layers[l];
                double* exw = exr + (size_t)(ex_buf * 4 * kFuExFields) * kI8Rows + cg * exJ;
                const double* exb = exr + (size_t)(ex_buf * 4 * kFuExFields) * kI8Rows;
                // the four workers of a row meet here: sum of the log-derivatives, rotation of the row vector
                // (x, the Householder components and the log-derivative are in the exchange already)
                auto meet = [&]() -> double {
                    FU_T(0);
                    bar_sync_named(2 + lq, 128);
                    FU_T(5);
                    double X[kFuMaxD], ld = 0.0;
#pragma unroll
                    for (int j = 0; j < kFuMaxD; ++j)
                        if (j < d) { X[j] = exb[j * exJ]; ld += exb[j * exJ + 5 * exF]; }
                    if (cg == 0) logdet_acc += DIR == JF_DIR_SAMPLE ? -ld : ld;
#pragma unroll 1
                    for (int ii = 0; ii < lc.hh_iter; ++ii) {
                        const int i = DIR == JF_DIR_LOGPDF ? ii : lc.hh_iter - 1 - ii;
                        double V[kFuMaxD], dot = 0, nrm = 0;
#pragma unroll
                        for (int j = 0; j < kFuMaxD; ++j)
                            if (j < d) {
                                V[j] = exb[j * exJ + (1 + i) * exF];
                                dot = fma(V[j], X[j], dot);
                                nrm = fma(V[j], V[j], nrm);
                            }
                        const double cc = 2.0 * dot / nrm;
#pragma unroll
                        for (int j = 0; j < kFuMaxD; ++j)
                            if (j < d) X[j] = fma(-cc, V[j], X[j]);
                    }
                    double mine = 0.0;
#pragma unroll
                    for (int j = 0; j < kFuMaxD; ++j) {
                        mine = (j == cg) ? X[j] : mine;
                        if (DIR == JF_DIR_SAMPLE && cg == 0 && j < d && !finite_(X[j])) n_bad |= 1;
                    }
                    if (G::kExBufs < 2) bar_sync_named(2 + lq, 128);        // single buffer: everybody has read before anybody writes again
                    return mine;
                };
                double v[SPT];
                if (DIR == JF_DIR_LOGPDF) {
                    static_assert(DIR != JF_DIR_LOGPDF || SPT == 12, "log_pdf layout: 12 slots per (tile, dimension)");
                    // tile 0: Householder components, offset, kernels 0-1
                    next_tile(v);
                    if (lc.has_offset) xj -= v[4];
                    exw[0] = xj;
#pragma unroll
                    for (int i = 0; i < 4; ++i) exw[(1 + i) * exF] = v[i];
                    exw[5 * exF] = logd_prev;
                    xj = meet();
                    FuSums S;
                    if (JF_FU_DBG & 2) { S.D = 0; S.E = 1; S.big_p = S.small_p = S.big_n = S.small_n = S.Sp = v[5]; S.ex = S.qc = 0; S.nsum = 1; S.n_pos = 1; }
                    else fu_stream<2, true>(S, lc, xj, v + 5);
                    // tiles 1, 2: four kernels each
#pragma unroll 1
                    for (int part = 1; part < 3; ++part) {
                        next_tile(v);
                        if (JF_FU_DBG & 2) S.Sp += v[0] + v[3] + v[7] + v[11];
                        else fu_stream<4, false>(S, lc, xj, v);
                    }
                    double y = xj, logd = 0.0;
                    if (active && live) {
                        const MixVal<double> mv = fu_stream_finish(S);
                        inv_stage(lc.inv_type, mv, y, logd);
                    }
                    xj = y;
                    logd_prev = logd;
                } else {
                    // slots of the sampling layout: w[0..9] | n[0..9] | mean[0..9] | v_0..3 | offset | padding.  The regulated
                    // kernels go to this worker's shared-memory slots ([field][k][worker], conflict free), where the root
                    // finder of csrc/gf.cuh (rolled loops over k, out of line) reads them.
                    double* slot_m = sSlots + tid, * slot_iw = slot_m + kFuK * kFuWorkers, * slot_n = slot_iw + kFuK * kFuWorkers;
                    double off = 0.0, nsum = 0.0, mmin = Num<double>::big, mmax = -Num<double>::big;
                    double hh[4] = {0.0, 0.0, 0.0, 0.0};        // this dimension's Householder components: exchanged after the solve
                    // slot s of this (layer, dimension): w[0..9] | n[0..9] | mean[0..9] | v_0..3 | offset | padding
                    auto take = [&](int s, double val) {
                        if (JF_FU_DBG & 2) { off += val; return; }
                        if (s < 10) slot_iw[s * kFuWorkers] = regulate_inv_width(val, lc.w_min, lc.inv_w_max);
                        else if (s < 20) {
                            const double g = regulate_norm(val, lc.n_min, lc.n_max);
                            nsum += g;
                            slot_n[(s - 10) * kFuWorkers] = g;
                        } else if (s < 30) {
                            slot_m[(s - 20) * kFuWorkers] = val;
                            mmin = tmin(mmin, val);
                            mmax = tmax(mmax, val);
                        } else if (s < 34) hh[s < 34 ? s - 30 : 0] = val;
                        else if (s == 34) off = val;
                    };
                    static_assert(DIR != JF_DIR_SAMPLE || (TPL - 1) * SPT >= 30, "the root finder's parameters must fit into the first TPL - 1 tiles");
#pragma unroll
                    for (int part = 0; part < TPL - 1; ++part) {
                        next_tile(v);
#pragma unroll
                        for (int i = 0; i < SPT; ++i) take(part * SPT + i, v[i]);
                    }
                    double logd = 0.0;
                    int ev = 0;
                    bool conv = true;
                    FU_T(0);
                    if (JF_FU_DBG & 2) xj += off;
                    else {
                        const double inv = 1.0 / nsum;
#pragma unroll
                        for (int k = 0; k < kFuK; ++k) slot_n[k * kFuWorkers] *= inv;
                        if (active && live) {
                            MixView<double> mv;
                            mv.m = smem_addr(slot_m); mv.iw = smem_addr(slot_iw); mv.n = smem_addr(slot_n);
                            mv.skb = (unsigned)(kFuWorkers * sizeof(double)); mv.K = kFuK; mv.mmin = mmin; mv.mmax = mmax;
                            xj = gf_solve<double>(mv, lc.inv_type, xj, logd, ev, conv);
                        }
                    }
                    FU_T(6);
                    n_evals += ev;
                    n_unconv += conv ? 0 : 1;
                    // the last tile of the layer (the remaining reflection components and the offset) was produced while
                    // this worker was solving
                    next_tile(v);
#pragma unroll
                    for (int i = 0; i < SPT; ++i) take((TPL - 1) * SPT + i, v[i]);
                    exw[0] = xj;                                // (own mean slots: the solve is over)
#pragma unroll
                    for (int i = 0; i < 4; ++i) exw[(1 + i) * exF] = hh[i];
                    exw[5 * exF] = logd;
                    xj = meet();
                    if (lc.has_offset) xj += off;
                }
                if (G::kExBufs == 2) ex_buf ^= 1;
            }
            // ---- outputs ----
            if (DIR == JF_DIR_LOGPDF) {
                // final exchange: base coordinates and the last layer's log-derivatives -> worker 0 of the row
                double* exw = exr + (size_t)(ex_buf * 4 * kFuExFields) * kI8Rows + cg * exJ;
                exw[0] = xj;
                exw[5 * exF] = logd_prev;
                bar_sync_named(2 + lq, 128);
                if (cg == 0) {
                    const double* exb = exr + (size_t)(ex_buf * 4 * kFuExFields) * kI8Rows;
                    double ld = 0.0, zsq_b = 0.0;
                    for (int j = 0; j < d; ++j) {
                        const double xb = exb[j * exJ];
                        ld += exb[j * exJ + 5 * exF];
                        zsq_b = fma(xb, xb, zsq_b);
                        if (!finite_(xb)) n_bad |= 1;
                    }
                    logdet_acc += ld;
                    zsq = zsq_b;
                }
                if (G::kExBufs == 2) ex_buf ^= 1;
            }
            if (live) {
                if (active) a.out[row * a.ld_out + cg] = xj;
                if (cg == 0) {
                    const double logdet = logdet_acc;
                    if (!finite_(logdet)) n_bad |= 1;
                    if (a.logdet_out) a.logdet_out[row] = logdet;
                    if (a.logbase_out) {
                        const double prev = a.logbase_in ? a.logbase_in[row] : 0.0;
                        a.logbase_out[row] = prev - 0.5 * zsq - (double)d * kLogSqrt2Pi;
                    }
                } else if (DIR == JF_DIR_SAMPLE && active && !finite_(xj)) {
                    n_bad |= 2;                     // (an offset made it non-finite: worker 0 saw a finite value)
                }
            }
            if (!live) { n_bad = 0; }
            if (n_bad) { status_add(a.status, JF_STATUS_NONFINITE, 1); n_bad = 0; }
        }
#if JF_FU_PROF
        FU_T(0);
        if (blockIdx.x == 0 && lane == 0 && lq == 0)
            printf("worker warp %d blocks %d cycles/block: compute %lld | wait tile %lld | gather %lld | bar512 %lld | prologue %lld | meet barrier %lld | solve %lld\n",
                   warp, my_blocks, prof_[0] / my_blocks, prof_[1] / my_blocks, prof_[2] / my_blocks, prof_[3] / my_blocks, prof_[4] / my_blocks,
                   prof_[5] / my_blocks, prof_[6] / my_blocks);
#endif
        if (DIR == JF_DIR_SAMPLE) {
            int tot = n_evals, unc = n_unconv;
            for (int o = 16; o > 0; o >>= 1) {
                tot += __shfl_down_sync(0xffffffffu, tot, o);
                unc += __shfl_down_sync(0xffffffffu, unc, o);
            }
            if (lane == 0) { status_add(a.status, JF_STATUS_ITERATIONS, tot); status_add(a.status, JF_STATUS_UNCONVERGED, unc); }
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

}  // namespace jf
