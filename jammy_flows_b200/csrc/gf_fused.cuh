// Fused parameter generator + "g" chain: one kernel computes the per-row flow parameters of a conditional
// Euclidean sub-pdf on the tensor cores AND consumes them in the layer chain, so the [P, rows] parameter block (4.4 KB
// per row for the README flow) never exists in HBM.
//
//   reference hand-off: main/default.py:956 (MLP call) -> :998-1029 (layer loop, `extra_inputs` slices), sampling
//   :1438 -> :1482-1506; layer math gaussianization_flow.py:389-1114 (same arithmetic as csrc/gf.cuh, register-resident).
//
// Structure (one persistent CTA per SM, 128 rows per block, 20 warps):
//   warps 0-15  workers (112 registers each, setmaxnreg).  Prologue: layer 1 + tanh + int8 digits of the hidden
//               activations -> A slices in shared memory (identical to mlp2_i8_kernel).  Then worker (row, j) = (TMEM
//               lane, column group) owns DIMENSION j of its row: per layer it drains ITS 36 parameter columns (3 MMA tiles
//               x 12 columns) straight from the TMEM accumulators into registers -- the level accumulators are combined
//               in 64-bit INTEGER arithmetic (the ALU pipe idles, the FP64 pipe is the bound) -- and consumes them:
//                 log_pdf   STREAMING: a tile carries (mean, log-width, log-norm) TRIPLES, each kernel is regulated and
//                           added to the mixture sums as it arrives, nothing but the sums stays live (tile 0 opens with the
//                           Householder components and the offset, so the rotated coordinate is known before the first
//                           kernel); an online rescaling exponent keeps the sums exact however far out x is;
//                 sampling  the K regulated kernels stay in registers for the root finder (csrc/gf.cuh's, register twin).
//               The four workers of a row meet once per layer through a 6-field shared-memory exchange for the rotation.
//   warp 16     one elected thread issues tcgen05.mma (kind::i8, M = 128, N = 48, K = 32) for tile t+1 as soon as the
//               workers have drained tile t; the MMAs of a tile overlap the workers' regulation / evaluation of what they
//               already hold -- the tensor pipe and the FP64 pipe run side by side.
//   warp 17     one thread streams the pre-sliced W2 tiles (L2 resident) with cp.async.bulk into a small ring.
//   warps 18-19 idle (they complete the service warpgroup: setmaxnreg is a warpgroup-wide instruction).
// W2's rows are permuted by the prep kernel into CONSUMPTION order (per direction): tile (c, part) holds, for the layer
// consumed c-th, the 12 columns of `part` for each of the 4 dimensions.
#pragma once
#include "mlp_i8.cuh"
#include "gf.cuh"
#include "gf_fused_launch.cuh"

namespace jf {

constexpr int kFuTN = 48;           // output columns per MMA tile = 4 column groups (dimensions) x 12
constexpr int kFuCG = 12;           // columns per (tile, dimension)
constexpr int kFuLvlStride = 64;    // TMEM columns between level accumulators
constexpr int kFuWorkers = 512;     // 16 warps
constexpr int kFuThreads = 640;     // + service warpgroup: MMA warp, producer warp, two idle warps
constexpr int kFuRegsWorker = 112;  // setmaxnreg moves registers inside the CTA's own allocation (640 x 96): 512 x 16 taken = 128 x 64 released
constexpr int kFuRegsService = 32;
constexpr int kFuExFields = 6;      // exchange: x, 4 Householder components, log-derivative
constexpr int kFuExBytes = 2 * kFuMaxD * kFuExFields * kI8Rows * 8;   // double buffered
constexpr int kFuPrivFields = 3;    // per-worker scratch in shared memory: offset, running logdet, sum of squares
constexpr int kFuPrivBytes = kFuMaxD * kFuPrivFields * kI8Rows * 8;
// the level accumulators of one output are combined as ONE 64-bit integer: sum_{l < 6} v_l 256^(5-l) < 2^62
// (|v_l| <= (l+1) 2^21); a 7th level is added in floating point
template <int NS> struct FuLv { static constexpr int kInt = NS < 6 ? NS : 6; };

template <int NS>
__host__ __device__ inline int64_t fu_prep_bytes(int n_layers) {
    const int64_t n_tiles = 3 * (int64_t)n_layers;
    return n_tiles * NS * kFuTN * kI8H + n_tiles * kFuTN * 16;
}

__host__ __device__ inline int fu_smem_bytes(int ns, int kin, int n_slots) {
    return ns * kI8Rows * kI8H + n_slots * kFuTN * kI8H + 512 + kFuExBytes + kFuPrivBytes + (kin + 1) * kI8H * 8 + kI8Rows * (kin | 1) * 8;
}

// source row of W2 / b2 (index into the sub-pdf's raw parameter vector) of fused column (tile, jj); -1: zero column.
// Slots of (layer, dimension j), 12 per part:
//   sampling  part 0: log_w[0..9], offset_j, log_n[0]      part 1: log_n[1..8], v_0[j] .. v_3[j]      part 2: mean[0..9], log_n[9], pad
//   log_pdf   part 0: v_0[j] .. v_3[j], offset_j, (mean, log_w, log_n)[0], [1], pad      part 1: triples 2..5      part 2: triples 6..9
__host__ __device__ inline int fu_source_param(const FuLayerC& c, int d, int direction, int part, int jj) {
    const int cg = jj / kFuCG, s = jj - cg * kFuCG;
    if (cg >= d) return -1;
    const int K = kFuK;
    const int off_hh = c.raw_off + (c.has_offset ? d : 0);
    const int off_m = off_hh + c.hh_iter * d, off_w = off_m + K * d, off_n = off_w + K * d;
    if (direction == JF_DIR_LOGPDF) {
        int t = s;                       // index into the triple stream of this part
        int k0 = 2 + 4 * (part - 1);
        if (part == 0) {
            if (s < 4) return s < c.hh_iter ? off_hh + s * d + cg : -1;
            if (s == 4) return c.has_offset ? c.raw_off + cg : -1;
            if (s == 11) return -1;
            t = s - 5;
            k0 = 0;
        }
        const int k = k0 + t / 3, f = t - 3 * (t / 3);
        return (f == 0 ? off_m : (f == 1 ? off_w : off_n)) + k * d + cg;
    }
    if (part == 0) {
        if (s < 10) return off_w + s * d + cg;
        if (s == 10) return c.has_offset ? c.raw_off + cg : -1;
        return off_n + cg;
    }
    if (part == 1) {
        if (s < 8) return off_n + (s + 1) * d + cg;
        const int i = s - 8;
        return i < c.hh_iter ? off_hh + i * d + cg : -1;
    }
    if (s < 10) return off_m + s * d + cg;
    if (s == 10) return off_n + 9 * d + cg;
    return -1;
}

// W2 [P,128] fp64 -> int8 slices of the fused tiles (UMMA K-major no-swizzle layout) + (scale, b2) per fused column
template <int NS>
__global__ void __launch_bounds__(128) fu_prep_kernel(const __grid_constant__ FuArgs a, const double* __restrict__ W2,
                                                      const double* __restrict__ b2, int direction, unsigned char* ws) {
    const int tile = blockIdx.x, n_tiles = gridDim.x;
    const int c = tile / 3, part = tile - 3 * c;
    const int l = direction == JF_DIR_LOGPDF ? a.n_layers - 1 - c : c;
    double2* cst = reinterpret_cast<double2*>(ws + (size_t)n_tiles * NS * kFuTN * kI8H);
    __shared__ double s_inv[kFuTN];
    __shared__ int s_src[kFuTN];
    for (int jj = threadIdx.x; jj < kFuTN; jj += blockDim.x) {
        const int src = fu_source_param(a.layers[l], a.d, direction, part, jj);
        double mx = 0.0;
        if (src >= 0)
            for (int k = 0; k < kI8H; ++k) mx = fmax(mx, fabs(W2[(size_t)src * kI8H + k]));
        int e = 0;
        if (mx > 0.0) { frexp(mx, &e); }
        s_inv[jj] = ldexp(1.0, -e);
        s_src[jj] = src;
        cst[tile * kFuTN + jj] = src >= 0 ? make_double2(ldexp(1.0, e - 12 - 8 * (FuLv<NS>::kInt - 1)), b2[src]) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    unsigned char* base = ws + (size_t)tile * NS * kFuTN * kI8H;
    for (int idx = threadIdx.x; idx < kFuTN * kI8H; idx += blockDim.x) {
        const int jj = idx / kI8H, k = idx - jj * kI8H;
        const int src = s_src[jj];
        const double w = src >= 0 ? W2[(size_t)src * kI8H + k] * s_inv[jj] : 0.0;
        const unsigned long long dg = to_digits<NS>(w);
        const int off = (k >> 4) * (kFuTN / 8) * 128 + (jj >> 3) * 128 + (jj & 7) * 16 + (k & 15);
#pragma unroll
        for (int s = 0; s < NS; ++s) base[(size_t)(NS - 1 - s) * kFuTN * kI8H + off] = (unsigned char)(dg >> (8 * s));
    }
}

JF_DEVINL void tmem_ld4(uint32_t taddr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
JF_DEVINL void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// 12 parameter values of this worker from the NS level accumulators: exact integer combine of the levels (ALU pipe),
// one conversion, one FMA with (scale, b2)
template <int NS>
JF_DEVINL void fu_drain(uint32_t tbase, const double2* __restrict__ cst, double* v) {
    constexpr int LI = FuLv<NS>::kInt;
#pragma unroll
    for (int c4 = 0; c4 < 3; ++c4) {
        int r[NS][4];
#pragma unroll
        for (int l = 0; l < NS; ++l) tmem_ld4(tbase + l * kFuLvlStride + c4 * 4, r[l]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            long long acc = r[0][j];
#pragma unroll
            for (int l = 1; l < LI; ++l) acc = acc * 256 + (long long)r[l][j];
            double sacc = __ll2double_rn(acc);
#pragma unroll
            for (int l = LI; l < NS; ++l) sacc = fma((double)r[l][j], 1.0 / (double)(1ull << (8 * (l - LI + 1))), sacc);
            const double2 sb = __ldg(cst + c4 * 4 + j);
            v[c4 * 4 + j] = fma(sacc, sb.x, sb.y);
        }
    }
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

template <bool FIRST>
JF_DEVINL void fu_stream(FuSums& A, const FuLayerC& lc, double x, double m, double w_raw, double n_raw) {
    const double iw = regulate_inv_width(w_raw, lc.w_min, lc.inv_w_max);
    const double n = regulate_norm(n_raw, lc.n_min, lc.n_max);
    const double a = (x - m) * iw, sa = fabs(a);
    if (FIRST) {
        A.D = sa > 512.0 ? sa : 0.0;
        A.E = 1.0;
        if (A.D > 0.0) A.E = exp_neg(-A.D);
        A.big_p = A.small_p = A.big_n = A.small_n = A.Sp = A.ex = A.qc = 0.0;
        A.nsum = n;
        A.n_pos = 0;
    } else {
        A.nsum += n;
        if (A.D > 0.0 && sa < A.D) {
            const double nd = sa > 512.0 ? sa : 0.0;
            const double f = exp_neg(nd - A.D);
            A.small_p *= f; A.small_n *= f; A.Sp *= f; A.qc *= f;
            A.D = nd;
            A.E = nd > 0.0 ? exp_neg(-nd) : 1.0;
        }
    }
    const double u = exp_neg(A.D - sa);
    const double e = u * A.E;
    const double rx = rcp_1to2(1.0 + e);
    const double nr = n * rx, nur = nr * u;
    const double pt = nur * iw * rx;
    if (a >= 0.0) { A.big_p += nr; A.small_p += nur; ++A.n_pos; }
    else          { A.big_n += nr; A.small_n += nur; }
    A.Sp += pt;
    if (a < -20.0) {                               // softplus-threshold quirk of the reference, see mix_eval
        const double nq = n * e * rx;
        A.ex += nq;
        A.qc = fma(nq, u, A.qc);
        A.Sp = fma(nq * u * iw, 1.0 + rx, A.Sp);
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

// ---- register-resident twins of mix_eval / presolve_f32 / solve_logit / solve_general (csrc/gf.cuh); same arithmetic ----
struct FuMix {
    double m[kFuK], iw[kFuK], n[kFuK];
    double mmin, mmax;
};

template <bool NEED_D>
JF_DEVINL MixVal<double> fu_mix_eval(const FuMix& p, double x) {
    const bool all_neg = x < p.mmin, all_pos = x > p.mmax;
    double delta = 0;
    if (all_neg || all_pos) {
        delta = Num<double>::big;
#pragma unroll
        for (int k = 0; k < kFuK; ++k) delta = tmin(delta, fabs((x - p.m[k]) * p.iw[k]));
    }
    const double E = (delta > 0.0) ? exp_neg(-delta) : 1.0;
    double big_p = 0, small_p = 0, big_n = 0, small_n = 0, Sp = 0, ex = 0, qc = 0, Sd = 0;
#pragma unroll
    for (int k = 0; k < kFuK; ++k) {
        const double iw = p.iw[k], n = p.n[k];
        const double a = (x - p.m[k]) * iw;
        const double u = exp_neg(delta - fabs(a));
        const double e = u * E;
        const double rx = rcp_1to2(1.0 + e);
        const double nr = n * rx, nur = nr * u;
        const double pt = nur * iw * rx;
        if (a >= 0.0) { big_p += nr; small_p += nur; }
        else          { big_n += nr; small_n += nur; }
        if (NEED_D) {
            const double dt = pt * iw * (rx - e * rx);
            Sd += (a >= 0.0) ? -dt : dt;
        }
        Sp += pt;
        if (a < -20.0) {
            const double nq = n * e * rx;
            ex += nq;
            qc = fma(nq, u, qc);
            Sp = fma(nq * u * iw, 1.0 + rx, Sp);
        }
    }
    MixVal<double> v;
    v.Sc = big_p + small_n + qc; v.Ss = small_p + big_n; v.Sp = Sp; v.ex = ex; v.Sd = Sd; v.E = E;
    v.dc = all_neg ? delta : 0.0;
    v.ds = all_pos ? delta : 0.0;
    v.dp = delta;
    return v;
}

JF_DEVINL double fu_presolve_f32(const FuMix& p, double t, double x0, double lo, double hi, double wmin) {
    const float tf = (float)t, lof = (float)lo, hif = (float)hi;
    const float stop = (float)(JF_PRE_STOP * wmin);
    float x = (float)x0;
#pragma unroll 1
    for (int it = 0; it < JF_PRE_ITERS; ++it) {
        float Sc = 0.f, Ss = 0.f, Sp = 0.f, Sd = 0.f;
#pragma unroll
        for (int k = 0; k < kFuK; ++k) {
            const float iw = (float)p.iw[k], n = (float)p.n[k];
            const float a = (x - (float)p.m[k]) * iw;
            const float e = ex2_approx(-fabsf(a) * 1.4426950408889634f);
            const float r = rcp_approx(1.f + e);
            const float nr = n * r, ner = nr * e;
            const bool pos = a >= 0.f;
            Sc += pos ? nr : ner;
            Ss += pos ? ner : nr;
            const float pt = ner * r * iw;
            Sp += pt;
            const float dt = pt * iw * (r - e * r);
            Sd += pos ? -dt : dt;
        }
        const float ics = rcp_approx(Sc * Ss);
        const float dy = Sp * ics;
        const float f = (lg2_approx(Sc) - lg2_approx(Ss)) * 0.6931471805599453f - tf;
        const float d2 = Sd * ics - dy * dy * (Ss - Sc);
        const float den = 2.f * dy * dy - f * d2;
        const float dx = (den > dy * dy) ? (2.f * f * dy * rcp_approx(den)) : (f * rcp_approx(dy));
        const float xn = x - dx;
        if (!(xn > lof && xn < hif)) break;
        x = xn;
        if (fabsf(dx) <= stop) break;
    }
    const double xr = (double)x;
    return (xr > lo && xr < hi) ? xr : x0;
}

JF_DEVINL LogitRoot<double> fu_solve_logit(const FuMix& p, double t, bool use_ex) {
    using T = double;
    T lo = Num<T>::big, hi = -Num<T>::big, x = 0, wmax = 0, wmin = Num<T>::big;
#pragma unroll
    for (int k = 0; k < kFuK; ++k) {
        const T m = p.m[k], w = rcp_pos(p.iw[k]);
        const T c = fma(t, w, m);
        lo = tmin(lo, c);
        hi = tmax(hi, c);
        x = fma(p.n[k], c, x);
        wmax = tmax(wmax, w);
        wmin = tmin(wmin, w);
    }
    {
        const T pad = T(1e-3) * wmax + T(64) * Num<T>::eps * (fabs(lo) + fabs(hi));
        lo -= pad;
        hi += pad;
        x = clampv(x, lo, hi);
    }
    if (fabs(t) < 60.0) x = fu_presolve_f32(p, t, x, lo, hi, wmin);
    const T tol_abs = Num<T>::newton_abs_tol, tol_rel = T(4) * Num<T>::eps;
    const T early = T(2e-6);
    LogitRoot<T> out;
    out.converged = false;
    out.evals = 0;
    out.x = x; out.logd = 0; out.lpdf = 0;
    T fprev = Num<T>::big, f = 0;
    const int kMaxIt = 64;
#pragma unroll 1
    for (int it = 0; it <= kMaxIt; ++it) {
        const MixVal<T> v = fu_mix_eval<true>(p, x);
        const T ssq = use_ex ? (v.Ss + v.ex) : v.Ss;
        if (it == kMaxIt) {          // iteration budget exhausted (non-finite parameters): report the last point
            out.x = x;
            out.logd = log(v.Sp / (v.Sc * ssq)) + (use_ex ? v.ex : T(0));
            out.lpdf = log(v.Sp) - v.dp;
            out.converged = false;
            return out;
        }
        ++out.evals;
        const T ics = rcp_pos(v.Sc * ssq);
        const T dy = v.Sp * ics;
        f = log(v.Sc * v.Sc * ics) + (v.ds - v.dc) - t;
        const T smc = (v.dc > T(0)) ? (ssq - v.Sc * v.E) : ((v.ds > T(0)) ? (ssq * v.E - v.Sc) : (ssq - v.Sc));
        const T d2 = v.Sd * ics - dy * dy * smc;
        const T den = T(2) * dy * dy - f * d2;
        const T dx = (den > dy * dy) ? (T(2) * f * dy * rcp_pos(den)) : (f * rcp_pos(dy));
        if (f < T(0)) lo = x; else hi = x;
        const T xn = x - dx;
        const bool inside = (xn > lo) && (xn < hi);
        const T adx = fabs(dx);
        const bool tiny = adx <= tol_abs + tol_rel * fabs(x);
        const bool at_noise = fabs(f) <= T(32) * Num<T>::eps * (T(1) + fabs(t));
        const bool small_step = inside && adx * dy <= early && adx <= early * wmin;
        if (small_step || tiny || at_noise) {
            const bool step = inside;
            out.x = step ? xn : x;
            const T sdx = step ? dx : T(0);
            if (use_ex) { out.logd = log(dy) + v.ex - sdx * d2 * rcp_pos(dy); out.lpdf = T(0); }
            else { out.lpdf = log(v.Sp) - v.dp - sdx * v.Sd * rcp_pos(v.Sp); out.logd = T(0); }
            out.converged = true;
            return out;
        }
        if (hi - lo <= tol_abs + tol_rel * fabs(x)) {
            out.x = x;
            if (use_ex) { out.logd = log(dy) + v.ex; out.lpdf = T(0); }
            else { out.lpdf = log(v.Sp) - v.dp; out.logd = T(0); }
            out.converged = fabs(f) <= Num<T>::target_prec;
            return out;
        }
        const bool shrinking = fabs(f) < T(0.75) * fprev;
        fprev = fabs(f);
        x = (inside && shrinking && finite_(xn)) ? xn : T(0.5) * (lo + hi);
    }
    return out;
}

// Pade tails of the inverse-normal variants (|z| > 5.32, ~1e-7 of all normals) and "inormal_full_pade": safeguarded
// Newton directly on y(x) = z, as solve_general in csrc/gf.cuh (cold)
JF_DEVINL double fu_solve_general(const FuMix& p, int type, double z, double& logd_out, int& evals, bool& converged) {
    using T = double;
    const T marg = (type == JF_INV_PARTLY_CRUDE ? T(0.6) : T(0.05)) + T(0.02) * fabs(z);
    const T t_lo = logit_phi<T>(z - marg), t_hi = logit_phi<T>(z + marg);
    T lo = Num<T>::big, hi = -Num<T>::big, x = 0, wmax = 0;
    const T t_mid = T(0.5) * (t_lo + t_hi);
#pragma unroll
    for (int k = 0; k < kFuK; ++k) {
        const T m = p.m[k], w = T(1) / p.iw[k];
        lo = tmin(lo, fma(t_lo, w, m));
        hi = tmax(hi, fma(t_hi, w, m));
        x = fma(p.n[k], fma(t_mid, w, m), x);
        wmax = tmax(wmax, w);
    }
    {
        const T pad = T(1e-3) * wmax + T(64) * Num<T>::eps * (fabs(lo) + fabs(hi));
        lo -= pad;
        hi += pad;
        x = clampv(x, lo, hi);
    }
    const T tol_abs = Num<T>::newton_abs_tol, tol_rel = T(4) * Num<T>::eps;
    T fprev = Num<T>::big, f = 0, logd = 0;
    converged = false;
    evals = 0;
    const int kMaxIt = 64;
#pragma unroll 1
    for (int it = 0; it < kMaxIt; ++it) {
        const MixVal<T> v = fu_mix_eval<true>(p, x);
        ++evals;
        T y;
        inv_stage(type, v, y, logd);          // log y' of the last evaluated point is what is reported
        f = y - z;
        if (f < T(0)) lo = x; else hi = x;
        const T dx = f / exp(logd);
        const T xn = x - dx;
        const bool inside = (xn > lo) && (xn < hi);
        if (fabs(dx) <= tol_abs + tol_rel * fabs(x)) {
            if (inside) x = xn;
            converged = true;
            break;
        }
        if (hi - lo <= tol_abs + tol_rel * fabs(x)) { converged = true; break; }
        const bool shrinking = fabs(f) < T(0.75) * fprev;
        fprev = fabs(f);
        x = (inside && shrinking && finite_(xn)) ? xn : T(0.5) * (lo + hi);
    }
    logd_out = logd;
    if (!(fabs(f) <= Num<T>::target_prec)) converged = false;
    return x;
}

JF_DEVINL double fu_solve(const FuMix& p, int type, double z, double& logd_out, int& evals, bool& converged) {
    if (type == JF_INV_ISIGMOID) {
        const LogitRoot<double> r = fu_solve_logit(p, z, true);
        logd_out = r.logd; evals = r.evals; converged = r.converged;
        return r.x;
    }
    if (type != JF_INV_FULL_PADE && fabs(z) < 5.32) {
        const LogitRoot<double> r = fu_solve_logit(p, logit_phi<double>(z), false);
        logd_out = kLogSqrt2Pi + 0.5 * z * z + r.lpdf;
        evals = r.evals; converged = r.converged;
        return r.x;
    }
    return fu_solve_general(p, type, z, logd_out, evals, converged);
}

// ---- main kernel ------------------------------------------------------------------------------------------------------
template <int REGS> JF_DEVINL void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> JF_DEVINL void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

template <int NS, int DIR, int KR>
__global__ void __launch_bounds__(kFuThreads, 1) gf_fused_kernel(const __grid_constant__ FuArgs a, int n_slots) {
    using Cfg = I8Cfg<NS, kFuTN>;
    extern __shared__ __align__(1024) unsigned char smem[];
    const MlpArgs<double>& m = a.m;
    const int Kin = m.dims[0];
    const int L = a.n_layers, n_tiles = 3 * L, d = a.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_blocks = (m.B + kI8Rows - 1) / kI8Rows;
    const int my_blocks = (int)((n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t offB = Cfg::offB, offBar = offB + n_slots * Cfg::kSliceBytesB;
    const uint32_t bar0 = sbase + offBar;
    // barriers (8 B each): full[16] | empty[16] | acc_full | acc_empty | a_full ; tmem pointer at +448
    auto bar_full = [&](int sl) { return bar0 + 8 * sl; };
    auto bar_empty = [&](int sl) { return bar0 + 8 * (16 + sl); };
    const uint32_t bar_acc_full = bar0 + 8 * 32, bar_acc_empty = bar0 + 8 * 33, bar_a_full = bar0 + 8 * 34;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + offBar + 448);
    double* sEx = reinterpret_cast<double*>(smem + offBar + 512);          // [2][4][6][128]
    double* sPriv = sEx + kFuExBytes / 8;                                  // [4][3][128]
    double* sW1 = sPriv + kFuPrivBytes / 8;                                // [Kin][128]
    double* sB1 = sW1 + (size_t)Kin * kI8H;
    double* sIn = sB1 + kI8H;                                              // [128][Kin|1]
    const int ldin = Kin | 1;

    if (warp == 17 && lane == 0) {
        for (int sl = 0; sl < n_slots; ++sl) { mbar_init(bar_full(sl), 1); mbar_init(bar_empty(sl), 1); }
        mbar_init(bar_acc_full, 1);
        mbar_init(bar_acc_empty, 16);
        mbar_init(bar_a_full, 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(bar0 + 448), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < Kin * kI8H; e += kFuThreads) {
        const int i = e / kI8H, u = e - i * kI8H;
        sW1[e] = m.wt[0][(size_t)u * Kin + i];
    }
    if (tid < kI8H) sB1[tid] = m.bias[0][tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kFuTN >> 3) << 17) | ((uint32_t)(kI8Rows >> 4) << 24);

    if (warp >= 16) {
        reg_dec<kFuRegsService>();
        if (warp == 17) {
            // ---- producer: stream every W2 slice of every block through the ring ----
            if (lane == 0) {
                const int loads_per_block = n_tiles * NS;
                const int64_t total = (int64_t)my_blocks * loads_per_block;
                int slot = 0, idx = 0;
                uint32_t wrap_par = 1;                      // parity to wait for on empty[slot]: first round passes
                for (int64_t i = 0; i < total; ++i) {
                    mbar_wait(bar_empty(slot), wrap_par);
                    mbar_expect_tx(bar_full(slot), Cfg::kSliceBytesB);
                    bulk_g2s(sbase + offB + slot * Cfg::kSliceBytesB, a.wsB + (size_t)idx * Cfg::kSliceBytesB, Cfg::kSliceBytesB,
                             bar_full(slot));
                    if (++slot == n_slots) { slot = 0; wrap_par ^= 1u; }
                    if (++idx == loads_per_block) idx = 0;
                }
            }
        } else if (warp == 16) {
            // ---- MMA issuer ----
            if (elect_one()) {
                const uint32_t desc_hi = (Cfg::kSbo >> 4) | (1u << 14);
                const uint32_t a_lo0 = (((sbase + Cfg::offA) & 0x3FFFF) >> 4) | ((uint32_t)(Cfg::kLboA >> 4) << 16);
                const uint32_t b_lo0 = (((sbase + offB) & 0x3FFFF) >> 4) | ((uint32_t)(Cfg::kLboB >> 4) << 16);
                int slot = 0;
                uint32_t full_par = 0, tile_par = 0;
                for (int jb = 0; jb < my_blocks; ++jb) {
                    mbar_wait(bar_a_full, (uint32_t)(jb & 1));        // the workers have written this block's A slices
                    tc_fence_after();
                    for (int t = 0; t < n_tiles; ++t, tile_par ^= 1u) {
                        mbar_wait(bar_acc_empty, tile_par ^ 1u);      // the workers have drained the previous tile
                        tc_fence_after();
#pragma unroll
                        for (int q = 0; q < NS; ++q) {
                            mbar_wait(bar_full(slot), full_par);
                            tc_fence_after();
                            const uint32_t b_lo = b_lo0 + slot * (Cfg::kSliceBytesB >> 4);
#pragma unroll
                            for (int p = 0; p + q < NS; ++p)
                                tc_mma_i8_x4(tmem + (p + q) * kFuLvlStride, a_lo0 + p * (Cfg::kSliceBytesA >> 4), b_lo, desc_hi,
                                             desc_hi, idesc, q > 0 ? 1u : 0u, (2 * Cfg::kLboA) >> 4, (2 * Cfg::kLboB) >> 4);
                            tc_commit(bar_empty(slot));
                            if (++slot == n_slots) { slot = 0; full_par ^= 1u; }
                        }
                        tc_commit(bar_acc_full);
                    }
                }
            }
        }
    } else {
        // ---- workers ----
        reg_inc<kFuRegsWorker>();
        const int lq = warp & 3, cg = warp >> 2;
        const int r = lq * 32 + lane;
        const bool active = cg < d;
        const uint32_t tbase = tmem + ((uint32_t)(lq * 32) << 16) + cg * kFuCG;
        uint32_t tile_par = 0;
        int ex_buf = 0;
        int n_evals = 0, n_unconv = 0, n_bad = 0;
        double* exr = sEx + r;                          // element (buf, j, f) at exr[((buf*4 + j)*6 + f)*128]
        // wait for the next tile, drain this worker's 12 columns, hand the accumulators back
        auto next_tile = [&](const double2* cst, double* v) {
            mbar_wait(bar_acc_full, tile_par); tile_par ^= 1u;
            tc_fence_after();
            fu_drain<NS>(tbase, cst, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty);
        };
#pragma unroll 1
        for (int jb = 0; jb < my_blocks; ++jb) {
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)jb * gridDim.x) * kI8Rows;
            const int64_t row = row0 + r;
            const bool live = row < m.B;
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
            bar_sync_named(1, kFuWorkers);
            // ---- prologue: layer 1 + tanh + digits -> A slices (thread = (row, quarter of the hidden units)) ----
            {
                const int quarter = cg;
                double in[KR];
#pragma unroll
                for (int i = 0; i < KR; ++i) in[i] = (i < Kin) ? sIn[r * ldin + i] : 0.0;
                const uint32_t a_row = sbase + Cfg::offA + (r >> 3) * 128 + (r & 7) * 16;
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
                            dg[uu] = to_digits<NS>(tanh_abs(z));
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
                    for (int sd = 0; sd < NS; ++sd) {
                        const uint32_t addr = a_row + (NS - 1 - sd) * Cfg::kSliceBytesA + chunk * Cfg::kLboA;
                        uint32_t w0, w1, w2, w3;
                        if (sd < 4) { w0 = lo[0][sd]; w1 = lo[1][sd]; w2 = lo[2][sd]; w3 = lo[3][sd]; }
                        else { w0 = hi[0][sd & 3]; w1 = hi[1][sd & 3]; w2 = hi[2][sd & 3]; w3 = hi[3][sd & 3]; }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a_full);

            // ---- the layer chain of this row, dimension cg ----
            double xj = (live && active) ? a.in[row * a.ld_in + cg] : 0.0;
            double* priv = sPriv + (size_t)(cg * kFuPrivFields) * kI8Rows + r;   // [0] offset, [1] logdet, [2] zsq
            if (cg == 0) {
                priv[kI8Rows] = (live && a.logdet_in) ? a.logdet_in[row] : 0.0;
                double zsq = 0.0;
                if (DIR == JF_DIR_SAMPLE && live)
                    for (int j = 0; j < d; ++j) { const double zj = a.in[row * a.ld_in + j]; zsq = fma(zj, zj, zsq); }
                priv[2 * kI8Rows] = zsq;
            }
            double logd_prev = 0.0;                     // log_pdf: log-derivative of the previous layer, not yet exchanged
#pragma unroll 1
            for (int c = 0; c < L; ++c) {
                const int l = DIR == JF_DIR_LOGPDF ? L - 1 - c : c;
                const FuLayerC& lc = a.layers[l];
                const double2* cst = a.consts + (size_t)(3 * c) * kFuTN + cg * kFuCG;
                double* exw = exr + (size_t)((ex_buf * 4 + cg) * kFuExFields) * kI8Rows;
                const double* exb = exr + (size_t)(ex_buf * 4 * kFuExFields) * kI8Rows;
                // the four workers of a row meet here: sum of the log-derivatives, rotation of the row vector
                // (x, the Householder components and the log-derivative are in the exchange already)
                auto meet = [&]() -> double {
                    bar_sync_named(2 + lq, 128);
                    double X[kFuMaxD], ld = 0.0;
#pragma unroll
                    for (int j = 0; j < kFuMaxD; ++j)
                        if (j < d) { X[j] = exb[(j * kFuExFields) * kI8Rows]; ld += exb[(j * kFuExFields + 5) * kI8Rows]; }
                    if (cg == 0) priv[kI8Rows] += DIR == JF_DIR_SAMPLE ? -ld : ld;
#pragma unroll 1
                    for (int ii = 0; ii < lc.hh_iter; ++ii) {
                        const int i = DIR == JF_DIR_LOGPDF ? ii : lc.hh_iter - 1 - ii;
                        double V[kFuMaxD], dot = 0, nrm = 0;
#pragma unroll
                        for (int j = 0; j < kFuMaxD; ++j)
                            if (j < d) {
                                V[j] = exb[(j * kFuExFields + 1 + i) * kI8Rows];
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
                    return mine;
                };
                double v[12];
                if (DIR == JF_DIR_LOGPDF) {
                    // tile 0: Householder components, offset, kernels 0-1
                    next_tile(cst, v);
                    if (lc.has_offset) xj -= v[4];
                    exw[0] = xj;
#pragma unroll
                    for (int i = 0; i < 4; ++i) exw[(1 + i) * kI8Rows] = v[i];
                    exw[5 * kI8Rows] = logd_prev;
                    xj = meet();
                    FuSums S;
                    fu_stream<true>(S, lc, xj, v[5], v[6], v[7]);
                    fu_stream<false>(S, lc, xj, v[8], v[9], v[10]);
                    // tiles 1, 2: four kernels each
#pragma unroll 1
                    for (int part = 1; part < 3; ++part) {
                        next_tile(cst + part * kFuTN, v);
#pragma unroll
                        for (int k = 0; k < 4; ++k) fu_stream<false>(S, lc, xj, v[3 * k], v[3 * k + 1], v[3 * k + 2]);
                    }
                    double y = xj, logd = 0.0;
                    if (active && live) {
                        const MixVal<double> mv = fu_stream_finish(S);
                        inv_stage(lc.inv_type, mv, y, logd);
                    }
                    xj = y;
                    logd_prev = logd;
                } else {
                    FuMix p;
                    // part 0: K log-widths, offset, first log-norm
                    next_tile(cst, v);
                    priv[0] = v[10];
#pragma unroll
                    for (int k = 0; k < kFuK; ++k) p.iw[k] = regulate_inv_width(v[k], lc.w_min, lc.inv_w_max);
                    p.n[0] = regulate_norm(v[11], lc.n_min, lc.n_max);
                    // part 1: 8 log-norms, this dimension's component of the Householder vectors
                    next_tile(cst + kFuTN, v);
#pragma unroll
                    for (int i = 0; i < 4; ++i) exw[(1 + i) * kI8Rows] = v[8 + i];
#pragma unroll
                    for (int k = 1; k < 9; ++k) p.n[k] = regulate_norm(v[k - 1], lc.n_min, lc.n_max);
                    // part 2: K means, last log-norm
                    next_tile(cst + 2 * kFuTN, v);
                    p.n[9] = regulate_norm(v[10], lc.n_min, lc.n_max);
                    {
                        double nsum = 0;
#pragma unroll
                        for (int k = 0; k < kFuK; ++k) nsum += p.n[k];
                        const double inv = 1.0 / nsum;
#pragma unroll
                        for (int k = 0; k < kFuK; ++k) p.n[k] *= inv;
                    }
                    p.mmin = Num<double>::big; p.mmax = -Num<double>::big;
#pragma unroll
                    for (int k = 0; k < kFuK; ++k) { p.m[k] = v[k]; p.mmin = tmin(p.mmin, v[k]); p.mmax = tmax(p.mmax, v[k]); }
                    double logd = 0.0;
                    int ev = 0;
                    bool conv = true;
                    if (active && live) xj = fu_solve(p, lc.inv_type, xj, logd, ev, conv);
                    n_evals += ev;
                    n_unconv += conv ? 0 : 1;
                    exw[0] = xj;
                    exw[5 * kI8Rows] = logd;
                    xj = meet();
                    if (lc.has_offset) xj += priv[0];
                }
                ex_buf ^= 1;
            }
            // ---- outputs ----
            if (DIR == JF_DIR_LOGPDF) {
                // final exchange: base coordinates and the last layer's log-derivatives -> worker 0 of the row
                double* exw = exr + (size_t)((ex_buf * 4 + cg) * kFuExFields) * kI8Rows;
                exw[0] = xj;
                exw[5 * kI8Rows] = logd_prev;
                bar_sync_named(2 + lq, 128);
                if (cg == 0) {
                    const double* exb = exr + (size_t)(ex_buf * 4 * kFuExFields) * kI8Rows;
                    double ld = 0.0, zsq = 0.0;
                    for (int j = 0; j < d; ++j) {
                        const double xb = exb[(j * kFuExFields) * kI8Rows];
                        ld += exb[(j * kFuExFields + 5) * kI8Rows];
                        zsq = fma(xb, xb, zsq);
                        if (!finite_(xb)) n_bad |= 1;
                    }
                    priv[kI8Rows] += ld;
                    priv[2 * kI8Rows] = zsq;
                }
                ex_buf ^= 1;
            }
            if (live) {
                if (active) a.out[row * a.ld_out + cg] = xj;
                if (cg == 0) {
                    const double logdet = priv[kI8Rows];
                    if (!finite_(logdet)) n_bad |= 1;
                    if (a.logdet_out) a.logdet_out[row] = logdet;
                    if (a.logbase_out) {
                        const double prev = a.logbase_in ? a.logbase_in[row] : 0.0;
                        a.logbase_out[row] = prev - 0.5 * priv[2 * kI8Rows] - (double)d * kLogSqrt2Pi;
                    }
                } else if (DIR == JF_DIR_SAMPLE && active && !finite_(xj)) {
                    n_bad |= 2;                     // (an offset made it non-finite: worker 0 saw a finite value)
                }
            }
            if (!live) { n_bad = 0; }
            if (n_bad) { status_add(a.status, JF_STATUS_NONFINITE, 1); n_bad = 0; }
        }
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
