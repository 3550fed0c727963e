// Gaussianization-flow layer "g": device math.
//
// What it computes is fixed by the reference (layers/euclidean/gaussianization_flow.py, cited per function); HOW is
// B200-first: one thread owns one row; the regulated mixture parameters of the current (layer, dimension) sit in
// shared memory (one CTA-wide table for shared parameters, per-thread conflict-free slots for per-row parameters); the
// logistic mixture is evaluated in LINEAR space with a single rescaling exponent (1 exp + 1 reciprocal per kernel
// instead of the reference's softplus + 3 logsumexp = 4 exp + 1 log1p per kernel); the sampling direction is a
// register-resident bracketed Newton iteration (no [B,K,d] temporaries, no host sync per iteration).
//
// Code size matters here: the first version unrolled every K loop and inlined every special function, which gave
// 200 KB of SASS per kernel and an instruction-fetch bound kernel (ncu: stall_no_instruction 5.8 per issue, see
// profiles/).  Loops over K are therefore rolled (unroll 2) and the cold inverse-CDF branches are not inlined.
#pragma once
#include "common.cuh"

// unroll factor of the loops over the K mixture kernels.  Measured on B200 (tools/build_variants.py + tools/kbench.py):
// rolled (1) beats 2 by 4-14 % and 4 by 25 % on the sampling kernels -- the kernels are instruction-cache bound
// (ncu stall_no_instruction), not latency bound (an Estrin-scheme exp with a 3x shorter dependency chain gains nothing)
#ifndef JF_K_UNROLL
#define JF_K_UNROLL 1
#endif
#ifndef JF_PREFETCH_NEXT_DIM
#define JF_PREFETCH_NEXT_DIM 1
#endif
#define JF_PRAGMA_(x) _Pragma(#x)
#define JF_PRAGMA(x) JF_PRAGMA_(x)
#define JF_UNROLL_K JF_PRAGMA(unroll JF_K_UNROLL)

namespace jf {

// 1/x for a positive normal x (hardware seed + two Newton steps; no special-case handling)
JF_DEVINL double rcp_pos_(double x) { return rcp_1to2(x); }
JF_DEVINL float rcp_pos_(float x) { return rcp_1to2(x); }   // rcp.approx.ftz: 1 ulp, no slow path

// ---------------------------------------------------------------------------------------------------------------------
// Layer constants (host fills from JfLayerDesc)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
struct GfLayerC {
    int K, d, hh_iter, inv_type, norm_mode, has_offset;
    int kind;      // 0: "g" gf_block, 1: "t" mvn_block (inv_type then holds the cov_type: 0 identity, 1 diagonal_symmetric, 2 diagonal, 3 full)
    int raw_off;   // start of the layer's slice in the raw parameter vector
    int tab_off;   // start of the layer's block in the processed shared-memory table
    T w_min, inv_w_max, n_min, n_max;
    // raw slice: [offset d][vs hh_iter*d][means K*d][log_w K*d][log_n K*d]   (index inside K*d blocks: k*d + j)
    __host__ __device__ int raw_hh() const { return raw_off + (has_offset ? d : 0); }
    __host__ __device__ int raw_m() const { return raw_hh() + hh_iter * d; }
    __host__ __device__ int raw_w() const { return raw_m() + K * d; }
    __host__ __device__ int raw_n() const { return raw_w() + K * d; }
    // processed table block: [offset d][vhat hh_iter*d][m K*d][w K*d][iw K*d][n K*d][mmin d][mmax d]
    __host__ __device__ int tab_hh() const { return tab_off + d; }
    __host__ __device__ int tab_m() const { return tab_hh() + hh_iter * d; }
    __host__ __device__ int tab_w() const { return tab_m() + K * d; }
    __host__ __device__ int tab_iw() const { return tab_w() + K * d; }
    __host__ __device__ int tab_n() const { return tab_iw() + K * d; }
    __host__ __device__ int tab_mmin() const { return tab_n() + K * d; }
    __host__ __device__ int tab_mmax() const { return tab_mmin() + d; }
    __host__ __device__ int tab_size() const { return kind == 1 ? 0 : d + hh_iter * d + 4 * K * d + 2 * d; }
};

// ---------------------------------------------------------------------------------------------------------------------
// Parameter regulation (reference gaussianization_flow.py:23-47, 300-317, 342, 406)
// ---------------------------------------------------------------------------------------------------------------------
// width:  w = w_min + 1/(1/w_max + exp(-raw))  ->  1/w = q/(w_min q + 1), q = 1/w_max + exp(-raw)   (one exp, one rcp)
template <typename T>
JF_DEVINL T regulate_inv_width(T raw, T w_min, T inv_w_max) {
    const T q = inv_w_max + exp_clamped(-raw);
    return q * rcp_pos_(fma(w_min, q, T(1)));
}
template <typename T>
JF_DEVINL void regulate_width(T raw, T w_min, T inv_w_max, T& w, T& iw) {
    iw = regulate_inv_width(raw, w_min, inv_w_max);
    w = T(1) / iw;
}
// norm (unnormalised, linear space): exp(log_n_regulated) = n_min + n_max * sigmoid(raw)
template <typename T>
JF_DEVINL T regulate_norm(T raw, T n_min, T n_max) {
    return fma(n_max, rcp_pos_(T(1) + exp_clamped(-raw)), n_min);
}

// View of the K regulated mixture parameters of one (layer, dimension).  Both homes of the parameters are shared
// memory, so the view holds 32-bit shared-window byte addresses and the loads are explicit ld.shared (a generic pointer
// would compile to LD with address-space resolution):
//   shared parameters: the CTA-wide table (stride d elements between consecutive k, broadcast reads);
//   per-row parameters: this thread's slots (stride blockDim.x elements, conflict-free reads).
template <typename T>
struct MixView {
    unsigned m, iw, n;   // byte addresses of element k = 0
    unsigned skb;        // byte stride between consecutive k
    int K;
    T mmin, mmax;        // hull of the means
};

JF_DEVINL double lds(unsigned addr, double) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
JF_DEVINL float lds(unsigned addr, float) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <typename T> JF_DEVINL T mv_m(const MixView<T>& v, int k) { return lds(v.m + k * v.skb, T()); }
template <typename T> JF_DEVINL T mv_iw(const MixView<T>& v, int k) { return lds(v.iw + k * v.skb, T()); }
template <typename T> JF_DEVINL T mv_n(const MixView<T>& v, int k) { return lds(v.n + k * v.skb, T()); }
JF_DEVINL unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Per-row parameters: read the raw values of dimension j (coalesced, param-major), regulate, write this thread's slots.
//   slots layout: field f in {m, iw, n}, element k at slots[(f*K + k)*T + tid]   (T = blockDim.x)
// NOTE: the asm loads above are not ordered against these C++ stores by the compiler; the slots are only read through
// mix_eval/gf_solve, which are separate (noinline) functions called after this one returns.
template <typename T>
__device__ __noinline__ MixView<T> regulate_to_slots(const GfLayerC<T>& c, int K, int j, const T* p, int64_t sj, T* slots,
                                                     bool prefetch_next_dim = true) {
    const int d = c.d, nt = blockDim.x;
    T* sm = slots + threadIdx.x;
    T* si = sm + (size_t)K * nt;
    T* sn = si + (size_t)K * nt;
    const T* pm = p + (int64_t)(c.raw_m() + j) * sj;
    const T* pw = p + (int64_t)(c.raw_w() + j) * sj;
    const T* pn = p + (int64_t)(c.raw_n() + j) * sj;
    const int64_t step = (int64_t)d * sj;
    // phase 1: raw values -> slots.  A pure copy loop keeps 3*8 independent global loads in flight per thread, so the
    // L2/HBM latency of the parameter stream is paid once per (layer, dim) and not once per k.
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        sm[(size_t)k * nt] = pm[k * step];
        si[(size_t)k * nt] = pw[k * step];
        sn[(size_t)k * nt] = (c.norm_mode != JF_NORM_NONE) ? pn[k * step] : T(0);
    }
#if JF_PREFETCH_NEXT_DIM
    // the parameters of the NEXT dimension of this layer are first touched ~1.5 us from now (after this dimension's
    // regulation and evaluation): ask L2 for them already, so that those loads find them there instead of in HBM
    // (the [P, rows] buffer is 2.3 GB per chunk, read exactly once: every first touch is an HBM access otherwise).
    // Measured: log_pdf per-row -9.6 %, sampling per-row -4.2 %; also prefetching across the layer boundary (next
    // layer's Householder vectors and first dimension) cost more instructions than it saved (+3 %) and was dropped.
    if (prefetch_next_dim && j + 1 < d) {
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pm + k * step + sj));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + k * step + sj));
            if (c.norm_mode != JF_NORM_NONE) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + k * step + sj));
        }
    }
#endif
    // phase 2: regulate in place (own slots only: no synchronisation needed)
    T nmax = -Num<T>::big;
    if (c.norm_mode == JF_NORM_RAW) {
        for (int k = 0; k < K; ++k) nmax = tmax(nmax, sn[(size_t)k * nt]);
    }
    T nsum = 0, mmin = Num<T>::big, mmax = -Num<T>::big;
    JF_UNROLL_K
    for (int k = 0; k < K; ++k) {
        const T m = sm[(size_t)k * nt];
        const T iw = regulate_inv_width(si[(size_t)k * nt], c.w_min, c.inv_w_max);
        T g;
        if (c.norm_mode == JF_NORM_REGULATED) g = regulate_norm(sn[(size_t)k * nt], c.n_min, c.n_max);
        else if (c.norm_mode == JF_NORM_RAW) g = exp(sn[(size_t)k * nt] - nmax);
        else g = T(1);
        nsum += g;
        mmin = tmin(mmin, m);
        mmax = tmax(mmax, m);
        si[(size_t)k * nt] = iw;
        sn[(size_t)k * nt] = g;
    }
    const T inv = T(1) / nsum;
    for (int k = 0; k < K; ++k) sn[(size_t)k * nt] *= inv;
    MixView<T> v;
    v.m = smem_addr(sm); v.iw = smem_addr(si); v.n = smem_addr(sn);
    v.skb = (unsigned)(nt * sizeof(T)); v.K = K; v.mmin = mmin; v.mmax = mmax;
    return v;
}

template <typename T>
JF_DEVINL MixView<T> table_view(const GfLayerC<T>& c, int K, int j, const T* tab) {
    MixView<T> v;
    v.m = smem_addr(tab + c.tab_m() + j); v.iw = smem_addr(tab + c.tab_iw() + j); v.n = smem_addr(tab + c.tab_n() + j);
    v.skb = (unsigned)(c.d * sizeof(T)); v.K = K;
    v.mmin = tab[c.tab_mmin() + j]; v.mmax = tab[c.tab_mmax() + j];
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// K-logistic mixture: log CDF / log SF / log PDF  (reference gaussianization_flow.py:389-454)
//   a_k = (x - m_k)/w_k ;  cdf = sum n_k sigma(a_k) ; sf = sum n_k sigma(-a_k) ; pdf = sum n_k sigma(a_k) sigma(-a_k)/w_k
// Linear-space evaluation with ONE shared rescaling exponent delta = min_k |a_k| when x lies outside the hull of the
// means (all a_k of one sign), so nothing underflows however far out x is:
//   u_k = exp(delta - |a_k|) in (0,1],  e_k = exp(-|a_k|) = u_k * exp(-delta),  r_k = 1/(1+e_k)
//   sigma(|a|) = r,  sigma(-|a|) = e r.   Results are returned as (S, shift) with log X = log S - shift.
//
// Reference quirk that is part of the numerical contract: F.softplus(t) returns t for t > 20 (torch default
// threshold), so for kernels with a_k < -20 the reference's log-terms drop the factor 1/(1+e^{a_k}):
//   cdf_k = n e^{a}        (exact n e^{a} r)      -- a 2e-9 relative change of a term that is itself <= 2e-9: invisible
//   sf_k  = n              (exact n r)            -- sf_ref = sf_exact + ex,  ex = sum_{a_k<-20} n_k e^{a_k} r_k <= 2e-9
//   pdf_k = n e^{a}/w      (exact n e^{a} r^2/w)
// so that cdf_ref + sf_ref = 1 + ex.  The kernel carries the exact sf (needed for an accurate Phi^-1 from the upper
// tail) and the excess separately.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
struct MixVal {
    T Sc, Ss, Sp;   // rescaled cdf / sf / pdf sums (Ss is the EXACT survival sum)
    T dc, ds, dp;   // shifts: log cdf = log Sc - dc, ...
    T ex;           // softplus-threshold excess of the reference's sf: sf_ref = Ss + ex
    T Sd;           // rescaled derivative of the pdf sum (same shift as Sp), for the third-order root step
    T E;            // exp(-delta)
};

// softplus-threshold quirk terms (see above) for the kernels with a_k < -20: rare (a kernel more than 20 widths to the
// right of x), so they are recomputed here, out of line, instead of riding along in the hot loop as predicated
// instructions.  The reference uses r = 1 instead of rx in the cdf, pdf and sf terms; q = 1 - rx = e*rx <= 2e-9.
#ifndef JF_QUIRK_OUTLINE
#define JF_QUIRK_OUTLINE 0
#endif
template <typename T>
__device__ __noinline__ void mix_quirk(const MixView<T>& mv, T x, T delta, T E, T& ex, T& qc, T& Sp) {
#pragma unroll 1
    for (int k = 0; k < mv.K; ++k) {
        const T iw = mv_iw(mv, k);
        const T a = (x - mv_m(mv, k)) * iw;
        if (a < T(-20)) {
            const T u = exp_neg(delta - fabs(a));
            const T e = u * E;
            const T rx = rcp_1to2(T(1) + e);
            const T nq = mv_n(mv, k) * e * rx;
            ex += nq;                                   // sf_ref - sf_exact
            qc = fma(nq, u, qc);                        // cdf_ref - cdf_exact (rescaled like small_n)
            Sp = fma(nq * u * iw, T(1) + rx, Sp);       // pdf_ref - pdf_exact
        }
    }
}

// x with the sign of -s (s != NaN): flips the sign bit of x when s >= 0
JF_DEVINL double neg_if_nonneg(double x, double s) {
    const int hs = __double2hiint(s);
    return __hiloint2double(__double2hiint(x) ^ (~hs & 0x80000000), __double2loint(x));
}
JF_DEVINL float neg_if_nonneg(float x, float s) { return s >= 0.f ? -x : x; }

// NEED_D: also accumulate the derivative of the pdf sum (only the sampling root finder uses it)
template <typename T, bool NEED_D = true>
JF_DEVINL MixVal<T> mix_eval(const MixView<T>& mv, T x) {
    const int K = mv.K;
    const bool all_neg = x < mv.mmin, all_pos = x > mv.mmax;   // 1/w > 0: sign(a_k) = sign(x - m_k)
    T delta = 0;
    if (all_neg || all_pos) {
        delta = Num<T>::big;
#pragma unroll 1
        for (int k = 0; k < K; ++k) delta = tmin(delta, fabs((x - mv_m(mv, k)) * mv_iw(mv, k)));
    }
    const T E = (delta > T(0)) ? exp_neg(-delta) : T(1);
    // four class sums instead of per-term selects: "big" = n*sigma(|a|), "small" = n*sigma(-|a|) (rescaled), by sign of a
    T big_p = 0, small_p = 0, big_n = 0, small_n = 0, Sp = 0, ex = 0, qc = 0, Sd = 0;
#if JF_QUIRK_OUTLINE
    T amin = 0;
#endif
    JF_UNROLL_K
    for (int k = 0; k < K; ++k) {
        const T iw = mv_iw(mv, k), n = mv_n(mv, k);
        const T a = (x - mv_m(mv, k)) * iw;
        const T u = exp_neg(delta - fabs(a));
        const T e = u * E;
        const T rx = rcp_1to2(T(1) + e);           // exact sigma(|a|)
        const T nr = n * rx, nur = nr * u;
        const T pt = nur * iw * rx;                 // pdf term n sigma(a) sigma(-a) / w
        if (a >= T(0)) { big_p += nr; small_p += nur; }
        else           { big_n += nr; small_n += nur; }
        if (NEED_D) {
            // d/dx of the pdf term: pt * (sigma(-a) - sigma(a)) / w = -+ pt * iw * rx * (1 - e)
#if JF_QUIRK_OUTLINE
            Sd += neg_if_nonneg(pt * iw * (rx - e * rx), a);
#else
            const T dt = pt * iw * (rx - e * rx);
            Sd += (a >= T(0)) ? -dt : dt;
#endif
        }
        Sp += pt;
#if JF_QUIRK_OUTLINE
        amin = tmin(amin, a);
#else
        // softplus-threshold quirk (a < -20, rare: a kernel more than 20 widths to the right of x): the reference uses
        // r = 1 instead of rx in the cdf and pdf terms and in the sf term; q = 1 - rx = e*rx <= 2e-9
        if (a < T(-20)) {
            const T nq = n * e * rx;
            ex += nq;                                   // sf_ref - sf_exact
            qc = fma(nq, u, qc);                        // cdf_ref - cdf_exact (rescaled like small_n)
            Sp = fma(nq * u * iw, T(1) + rx, Sp);       // pdf_ref - pdf_exact
        }
#endif
    }
#if JF_QUIRK_OUTLINE
    if (amin < T(-20)) mix_quirk<T>(mv, x, delta, E, ex, qc, Sp);
#endif
    const T Sc = big_p + small_n + qc;
    const T Ss = small_p + big_n;
    MixVal<T> v;
    v.Sc = Sc; v.Ss = Ss; v.Sp = Sp; v.ex = ex; v.Sd = Sd; v.E = E;
    v.dc = all_neg ? delta : T(0);
    v.ds = all_pos ? delta : T(0);
    v.dp = delta;
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// inverse-CDF stage and its log-derivative (reference gaussianization_flow.py:480-560 and :568-671)
// ---------------------------------------------------------------------------------------------------------------------
// the three inverse-normal variants: cold (only layer 0 of a sub-pdf uses them), kept out of line
template <typename T>
__device__ __noinline__ void inv_stage_inormal(int type, T lc, T ls, T lp, T sfx, T& y, T& logd) {
    const T eps = T(0.5e-7), pa = T(0.147);
    const T pc = T(2.0 / (kPi * 0.147));
    const T cdf = exp(lc);
    // upper bulk limit tested on the exact survival function (cdf < 1-eps <=> sf > eps): in fp32 1-eps rounds to 1 and
    // a cdf that rounds to 1-ulp would otherwise enter the bulk branch with an underflowed tail
    const bool upper = !(sfx > eps);   // reference: cdf >= 1 - eps
    const bool bulk = (cdf > eps) && !upper;
    if (type != JF_INV_FULL_PADE && bulk) {
        // Phi^-1(cdf) = sqrt2*erfinv(2cdf-1) in the reference (torch Normal.icdf); evaluated here from the SMALLER tail
        // with erfcinv so that the argument keeps full relative precision (2cdf-1 cancels catastrophically near 1).
        const T e = (cdf <= T(0.5)) ? -erfcinv(T(2) * cdf) : erfcinv(T(2) * sfx);
        y = T(1.4142135623730951) * e;
        logd = T(kLogSqrt2Pi) + e * e + lp;
        return;
    }
    if (type == JF_INV_PARTLY_CRUDE) {
        const T s = -T(2) * (ls + lc);
        const T tail = sqrt(s) - T(0.4717);
        y = upper ? tail : -tail;
        logd = -T(0.5) * log(s) - ls - lc + lp;
        return;
    }
    // Pade branch (tails of "partly_precise", everything for "full_pade")
    const T lnf = lc + ls + T(1.3862943611198906);   // log 4
    const T F = pc + lnf * T(0.5);
    const T F2 = sqrt(F * F - lnf / pa);
    const T pade = sqrt(tmax(T(0), T(2) * (F2 - F)));
    T total;
    if (cdf > T(0.49999) && cdf < T(0.50001)) {
        total = T(0.91893848994417);   // log(2.506628), reference :623-625 / :654
    } else {
        const T lnum = log(-(F - T(1) / pa - F2));
        const T lden = T(1.0397207708399179) + T(0.5) * log(F2 - F) + log(F2);   // 0.5*log 8
        total = lnum - lden - ls - lc + log(fabs(T(1) - T(2) * cdf));
    }
    if (type == JF_INV_FULL_PADE) y = (cdf <= T(0.5)) ? -pade : pade;
    else y = upper ? pade : -pade;
    logd = total + lp;
}

template <typename T>
JF_DEVINL void inv_stage(int type, const MixVal<T>& v, T& y, T& logd) {
    if (type == JF_INV_ISIGMOID) {
        // y = log cdf - log sf_ref ; logaddexp(-ls,-lc) + lp = lp - lc - ls + log(cdf+sf), cdf_ref + sf_ref = 1 + ex
        const T ssq = v.Ss + v.ex;
        y = log(v.Sc / ssq) + (v.ds - v.dc);
        logd = log(v.Sp / (v.Sc * ssq)) + v.ex;     // the shifts cancel exactly (dp = dc + ds)
        return;
    }
    const T lc = log(v.Sc) - v.dc, ls = log(v.Ss + v.ex) - v.ds, lp = log(v.Sp) - v.dp;
    inv_stage_inormal<T>(type, lc, ls, lp, v.Ss * exp(-v.ds), y, logd);
}

// one element of the log_pdf direction: y and log dy/dx (kept out of line: one copy per kernel, not one per dimension)
template <typename T>
__device__ __noinline__ void gf_eval_logpdf(const MixView<T>& mv, int type, T x, T& y, T& logd) {
    const MixVal<T> v = mix_eval<T, false>(mv, x);
    inv_stage(type, v, y, logd);
}

// log(Phi(t)/Phi(-t)): the logistic-scale target that corresponds to a Gaussian-scale target t
template <typename T>
__device__ __noinline__ T logit_phi(T t) {
    const T at = fabs(t);
    const T lim = sizeof(T) == 8 ? T(25) : T(12);
    if (at < lim) {
        const T r = T(0.70710678118654752);
        return log(erfc(-t * r) / erfc(t * r));
    }
    const T v = T(0.5) * at * at + log(at) + T(kLogSqrt2Pi);
    return t > 0 ? v : -v;
}

template <typename T> JF_DEVINL T rcp_pos(T x) { return rcp_pos_(x); }

// ---------------------------------------------------------------------------------------------------------------------
// Root finding for the sampling direction (reference: 25 bisections on [-1e5,1e5] + <=20 masked Newton steps,
// gaussianization_flow.py:921 + bisection_n_newton.py:11-135).  Same root, different trajectory.
//
// Every bulk root is found in LOGIT space: y(x) = z with y = logit(cdf) (isigmoid) or y = Phi^-1(cdf) (inverse-normal
// variants inside their bulk region, |z| < Phi^-1(1-0.5e-7) = 5.3267) is the root of
//        L(x) = log(cdf(x)/sf(x)) = t,      t = z  or  t = log(Phi(z)/Phi(-z)),
// so the iteration never evaluates erfinv or the Pade branches.
//   * analytic bracket: every kernel satisfies sigma((x-m_k)/w_k) <= sigma(t) for x <= min_k(m_k + t w_k) and
//     >= sigma(t) for x >= max_k(m_k + t w_k), so the mixture root lies in between;
//   * third-order (Halley) steps from sum_k n_k (m_k + t w_k), safeguarded by the bracket (bisect when a step leaves
//     it or |f| does not shrink);
//   * a step shorter than 1e-7 of the local scale lands within ~1e-21 of the root: it is accepted WITHOUT a confirming
//     evaluation, and log y' / log pdf are carried to the new point to first order (error ~1e-14).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
struct LogitRoot {
    T x;          // root
    T logd;       // log dL/dx at the root (+ the reference's softplus excess when use_ex)
    T lpdf;       // log pdf at the root
    int evals;
    bool converged;
};

// fp32 pre-solve of L(x) = t (sampling direction, fp64 kernels only): up to three Halley steps with the mixture evaluated
// in single precision on the FP32 / MUFU pipes (ex2.approx, rcp.approx, lg2.approx), which idle while the FP64 pipe is
// the bound of this kernel.  Its result is only the STARTING POINT of the fp64 iteration below -- accuracy and the
// acceptance test live entirely there -- but it lands within ~1e-6 widths of the root, so the fp64 loop needs one or
// two mixture evaluations instead of three or four (measured, profiles/).  Anything non-finite falls back to x0.
#ifndef JF_PRESOLVE_F32
#define JF_PRESOLVE_F32 1
#endif
JF_DEVINL float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
JF_DEVINL float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
JF_DEVINL float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#ifndef JF_PRE_CVT
#define JF_PRE_CVT 0      // 0: F2F.F32.F64 of the full double; 1: integer re-bias of the high word (20 mantissa bits)
#endif
#ifndef JF_PRE_ITERS
#define JF_PRE_ITERS 3
#endif
#ifndef JF_PRE_STOP
#define JF_PRE_STOP 4e-3
#endif
#ifndef JF_PRE_BRACKET32
#define JF_PRE_BRACKET32 0
#endif
// shared-memory double -> float for the pre-solve
JF_DEVINL float lds_as_f32(unsigned addr) {
#if JF_PRE_CVT
    unsigned h;
    asm("ld.shared.u32 %0, [%1+4];" : "=r"(h) : "r"(addr));
    const unsigned ex = h & 0x7ff00000u;
    const unsigned mag = (ex >= 0x38100000u && ex <= 0x47e00000u) ? (((h << 3) ^ 0x40000000u) & 0x7fffffffu) : 0u;
    return __uint_as_float(mag | (h & 0x80000000u));
#else
    return (float)lds(addr, double());
#endif
}

static __device__ __noinline__ double presolve_f32(const MixView<double>& mv, double t, double x0, double lo, double hi, double wmin) {
    const int K = mv.K;
    const float tf = (float)t, lof = (float)lo, hif = (float)hi;
    // a third-order step of length dx leaves an error ~dx^3/w^2: below JF_PRE_STOP widths the next point is already at
    // the single-precision floor and a confirming evaluation would be wasted
    const float stop = (float)(JF_PRE_STOP * wmin);
    float x = (float)x0;
#pragma unroll 1
    for (int it = 0; it < JF_PRE_ITERS; ++it) {
        float Sc = 0.f, Ss = 0.f, Sp = 0.f, Sd = 0.f;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const float iw = lds_as_f32(mv.iw + k * mv.skb), n = lds_as_f32(mv.n + k * mv.skb);
            const float a = (x - lds_as_f32(mv.m + k * mv.skb)) * iw;
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
        if (!(xn > lof && xn < hif)) break;      // also catches NaN: keep the last good point
        x = xn;
        if (fabsf(dx) <= stop) break;
    }
    const double xr = (double)x;
    return (xr > lo && xr < hi) ? xr : x0;
}

template <typename T> JF_DEVINL T presolve(const MixView<T>&, T, T x0, T, T, T) { return x0; }
#if JF_PRESOLVE_F32
template <> JF_DEVINL double presolve<double>(const MixView<double>& mv, double t, double x0, double lo, double hi, double wmin) {
    // the fp32 stage has no rescaling exponent: only where exp(-|t|) and the tails stay inside the fp32 range
    return (fabs(t) < 60.0) ? presolve_f32(mv, t, x0, lo, hi, wmin) : x0;
}
#endif

template <typename T>
__device__ __noinline__ LogitRoot<T> solve_logit(const MixView<T>& mv, T t, bool use_ex) {
    T lo = Num<T>::big, hi = -Num<T>::big, x = 0, wmax = 0, wmin = Num<T>::big;
    T bracket_eps = T(64) * Num<T>::eps;
#if JF_PRESOLVE_F32 && JF_PRE_BRACKET32
    if constexpr (sizeof(T) == 8) {
        // the bracket only safeguards the iteration: single precision (padded accordingly) is enough
        float lof = 1e30f, hif = -1e30f, xf = 0.f, wmaxf = 0.f, wminf = 1e30f;
        const float tf = (float)t;
        for (int k = 0; k < mv.K; ++k) {
            const float m = lds_as_f32(mv.m + k * mv.skb), w = rcp_approx(lds_as_f32(mv.iw + k * mv.skb));
            const float c = fmaf(tf, w, m);
            lof = fminf(lof, c);
            hif = fmaxf(hif, c);
            xf = fmaf(lds_as_f32(mv.n + k * mv.skb), c, xf);
            wmaxf = fmaxf(wmaxf, w);
            wminf = fminf(wminf, w);
        }
        lo = lof; hi = hif; x = xf; wmax = wmaxf; wmin = wminf;
        bracket_eps = T(1e-5);
    } else
#endif
#pragma unroll 1
    for (int k = 0; k < mv.K; ++k) {
        const T m = mv_m(mv, k), w = rcp_pos(mv_iw(mv, k));
        const T c = fma(t, w, m);
        lo = tmin(lo, c);
        hi = tmax(hi, c);
        x = fma(mv_n(mv, k), c, x);
        wmax = tmax(wmax, w);
        wmin = tmin(wmin, w);
    }
    {
        const T pad = T(1e-3) * wmax + bracket_eps * (fabs(lo) + fabs(hi));
        // (not clipped to the reference's bisection interval [-1e5,1e5]: its Newton steps leave that interval
        //  too, e.g. for logit targets ~1e3 behind a non-orthogonal triangular_combination "rotation")
        lo -= pad;
        hi += pad;
        x = clampv(x, lo, hi);
    }
    x = presolve<T>(mv, t, x, lo, hi, wmin);
    const T tol_abs = Num<T>::newton_abs_tol, tol_rel = T(4) * Num<T>::eps;
    // a third-order step of relative length s (in widths) lands within ~s^3 of the root, and log y' / log pdf carried to
    // first order are off by ~s^2: 2e-6 -> 1e-17 / 4e-12 (fp64).  The fp32 pre-solve typically leaves s ~ 1e-6.
    const T early = (sizeof(T) == 8 ? T(JF_PRESOLVE_F32 ? 2e-6 : 1e-7) : T(1e-3));
    LogitRoot<T> out;
    out.converged = false;
    out.evals = 0;
    T fprev = Num<T>::big, f = 0;
    const int kMaxIt = 64;
#pragma unroll 1
    for (int it = 0; it < kMaxIt; ++it) {
        const MixVal<T> v = mix_eval<T>(mv, x);
        ++out.evals;
        const T ssq = use_ex ? (v.Ss + v.ex) : v.Ss;
        const T ics = rcp_pos(v.Sc * ssq);
        const T dy = v.Sp * ics;                              // L'(x): the rescaling shifts cancel (dp = dc + ds)
        f = log(v.Sc * v.Sc * ics) + (v.ds - v.dc) - t;       // log(Sc/ssq)
        // L'' = p'/(c s) - (p/(c s))^2 (s - c) with the true (unscaled) s - c
        const T smc = (v.dc > T(0)) ? (ssq - v.Sc * v.E) : ((v.ds > T(0)) ? (ssq * v.E - v.Sc) : (ssq - v.Sc));
        const T d2 = v.Sd * ics - dy * dy * smc;
        const T den = T(2) * dy * dy - f * d2;                // Halley: dx = 2 f L' / (2 L'^2 - f L'')
        const T dx = (den > dy * dy) ? (T(2) * f * dy * rcp_pos(den)) : (f * rcp_pos(dy));
        if (f < T(0)) lo = x; else hi = x;
        const T xn = x - dx;
        const bool inside = (xn > lo) && (xn < hi);
        const T adx = fabs(dx);
        const bool tiny = adx <= tol_abs + tol_rel * fabs(x);
        // residual at the rounding level of its own evaluation: nothing left to gain (matters in fp32, where the noise of
        // f sits above the step-length tests and the iteration otherwise falls back to bisecting a wide bracket:
        // measured 22 evaluations per element on cfg4 before this test)
        const bool at_noise = fabs(f) <= T(32) * Num<T>::eps * (T(1) + fabs(t));
        // fp32: the step in logit units alone (the guard against the narrowest kernel is below the fp32 noise floor)
        const bool small_step = inside && adx * dy <= early && (sizeof(T) == 4 || adx <= early * wmin);
        if (small_step || tiny || at_noise) {
            const bool step = inside;                          // tiny but outside the bracket: stay
            out.x = step ? xn : x;
            const T sdx = step ? dx : T(0);
            // the callers use exactly one of the two: log L' for isigmoid (use_ex), log pdf for the inverse-normal stages
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
    // iteration budget exhausted (non-finite parameters): report the last point
    {
        const MixVal<T> v = mix_eval<T>(mv, x);
        const T ssq = use_ex ? (v.Ss + v.ex) : v.Ss;
        out.x = x;
        out.logd = log(v.Sp / (v.Sc * ssq)) + (use_ex ? v.ex : T(0));
        out.lpdf = log(v.Sp) - v.dp;
        out.converged = false;
    }
    return out;
}

// General (slow) path: safeguarded Newton directly on y(x) = z with a confirming evaluation.  Only used for targets in
// the Pade tails of the inverse-normal variants (|z| > 5.32, ~1e-7 of all normals) and for "inormal_full_pade".
template <typename T>
__device__ __noinline__ T solve_general(const MixView<T>& mv, int type, T z, T& logd_out, int& evals, bool& converged) {
    const T marg = (type == JF_INV_PARTLY_CRUDE ? T(0.6) : T(0.05)) + T(0.02) * fabs(z);
    const T t_lo = logit_phi<T>(z - marg), t_hi = logit_phi<T>(z + marg);
    T lo = Num<T>::big, hi = -Num<T>::big, x = 0, wmax = 0;
    const T t_mid = T(0.5) * (t_lo + t_hi);
    for (int k = 0; k < mv.K; ++k) {
        const T m = mv_m(mv, k), w = T(1) / mv_iw(mv, k);
        lo = tmin(lo, fma(t_lo, w, m));
        hi = tmax(hi, fma(t_hi, w, m));
        x = fma(mv_n(mv, k), fma(t_mid, w, m), x);
        wmax = tmax(wmax, w);
    }
    {
        const T pad = T(1e-3) * wmax + T(64) * Num<T>::eps * (fabs(lo) + fabs(hi));
        lo -= pad;
        hi += pad;
        x = clampv(x, lo, hi);
    }
    const T tol_abs = Num<T>::newton_abs_tol, tol_rel = T(4) * Num<T>::eps;
    T fprev = Num<T>::big, f = 0;
    converged = false;
    evals = 0;
    MixVal<T> v;
    const int kMaxIt = 64;
#pragma unroll 1
    for (int it = 0; it < kMaxIt; ++it) {
        v = mix_eval<T>(mv, x);
        ++evals;
        T y, logd;
        inv_stage(type, v, y, logd);
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
    T y;
    inv_stage(type, v, y, logd_out);
    if (!(fabs(f) <= Num<T>::target_prec)) converged = false;   // reference: warn above 1e-7 (fp64) / 1e-4 (fp32)
    return x;
}

// Root of y(x) = z for one element: returns x, log y'(x), the number of mixture evaluations and the convergence flag.
template <typename T>
JF_DEVINL T gf_solve(const MixView<T>& mv, int type, T z, T& logd_out, int& evals, bool& converged) {
    if (type == JF_INV_ISIGMOID) {
        const LogitRoot<T> r = solve_logit<T>(mv, z, true);
        logd_out = r.logd; evals = r.evals; converged = r.converged;
        return r.x;
    }
    // bulk region of the inverse-normal variants: Phi^-1(cdf) = z  <=>  logit(cdf) = logit(Phi(z))
    if (type != JF_INV_FULL_PADE && fabs(z) < T(5.32)) {
        const LogitRoot<T> r = solve_logit<T>(mv, logit_phi<T>(z), false);
        // log y' = log sqrt(2 pi) + erfinv(2cdf-1)^2 + log pdf with erfinv(2cdf-1) = z/sqrt2 at the root
        logd_out = T(kLogSqrt2Pi) + T(0.5) * z * z + r.lpdf;
        evals = r.evals; converged = r.converged;
        return r.x;
    }
    return solve_general<T>(mv, type, z, logd_out, evals, converged);
}

// ---------------------------------------------------------------------------------------------------------------------
// Householder rotation applied to the row vector (reference gaussianization_flow.py:457-471, :1038, :975)
//   Q = H_0 H_1 ... H_{n-1},  H_i = I - 2 v_i v_i^T/|v_i|^2.
//   log_pdf direction uses Q^T x = H_{n-1}(...H_0 x);  sampling uses Q x = H_0(...H_{n-1} x).
//   2 d FMAs per reflection on the vector instead of forming Q (d^3) as the reference does.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int DM>
JF_DEVINL void householder_apply(T* x, int d, int n_iter, bool transpose, bool processed, const T* vtab,
                                 const T* p, int64_t sj) {
#pragma unroll 1
    for (int ii = 0; ii < n_iter; ++ii) {
        const int i = transpose ? ii : (n_iter - 1 - ii);
        T v[DM];
        T dot = 0, nrm = 0;
        if (processed) {
#pragma unroll
            for (int j = 0; j < d; ++j) {
                v[j] = vtab[i * d + j];
                dot = fma(v[j], x[j], dot);
            }
            nrm = T(1);
        } else {
#pragma unroll
            for (int j = 0; j < d; ++j) {
                v[j] = p[(int64_t)(i * d + j) * sj];
                dot = fma(v[j], x[j], dot);
                nrm = fma(v[j], v[j], nrm);
            }
        }
        const T c = T(2) * dot / nrm;
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] = fma(-c, v[j], x[j]);
    }
}

// Build the processed shared-memory table of one layer from a raw SHARED parameter vector (one CTA, cooperative).
template <typename T>
__device__ void gf_build_table(const GfLayerC<T>& c, const T* raw /*stride 1*/, T* tab, int tid, int nthreads) {
    const int d = c.d, K = c.K;
    for (int j = tid; j < d; j += nthreads) tab[c.tab_off + j] = c.has_offset ? raw[c.raw_off + j] : T(0);
    for (int i = tid; i < c.hh_iter; i += nthreads) {
        T nrm = 0;
        for (int j = 0; j < d; ++j) { T v = raw[c.raw_hh() + i * d + j]; nrm = fma(v, v, nrm); }
        const T inv = T(1) / sqrt(nrm);
        for (int j = 0; j < d; ++j) tab[c.tab_hh() + i * d + j] = raw[c.raw_hh() + i * d + j] * inv;
    }
    for (int e = tid; e < K * d; e += nthreads) {
        tab[c.tab_m() + e] = raw[c.raw_m() + e];
        T w, iw;
        regulate_width(raw[c.raw_w() + e], c.w_min, c.inv_w_max, w, iw);
        tab[c.tab_w() + e] = w;
        tab[c.tab_iw() + e] = iw;
    }
    // normalisation and the hull of the means need a pass over k for each dimension
    for (int j = tid; j < d; j += nthreads) {
        T nmax = -Num<T>::big;
        if (c.norm_mode == JF_NORM_RAW)
            for (int k = 0; k < K; ++k) nmax = tmax(nmax, raw[c.raw_n() + k * d + j]);
        T sum = 0, mmin = Num<T>::big, mmax = -Num<T>::big;
        for (int k = 0; k < K; ++k) {
            T g;
            if (c.norm_mode == JF_NORM_REGULATED) g = regulate_norm(raw[c.raw_n() + k * d + j], c.n_min, c.n_max);
            else if (c.norm_mode == JF_NORM_RAW) g = exp(raw[c.raw_n() + k * d + j] - nmax);
            else g = T(1);
            tab[c.tab_n() + k * d + j] = g;
            sum += g;
            mmin = tmin(mmin, raw[c.raw_m() + k * d + j]);
            mmax = tmax(mmax, raw[c.raw_m() + k * d + j]);
        }
        const T inv = T(1) / sum;
        for (int k = 0; k < K; ++k) tab[c.tab_n() + k * d + j] *= inv;
        tab[c.tab_mmin() + j] = mmin;
        tab[c.tab_mmax() + j] = mmax;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One layer on one row.  `p`: this row's raw parameter slice base (element i at p[i*sj]); `tab`: processed table;
// `slots`: per-thread parameter slots (per-row mode).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int DM>
JF_DEVINL void gf_layer_logpdf(T* x, T& logdet, const GfLayerC<T>& c, int d, int K, bool processed, const T* tab,
                               const T* p, int64_t sj, T* slots) {
    if (c.has_offset) {
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] -= processed ? tab[c.tab_off + j] : p[(int64_t)(c.raw_off + j) * sj];
    }
    if (c.hh_iter > 0)
        householder_apply<T, DM>(x, d, c.hh_iter, true, processed, tab + c.tab_hh(), p + (int64_t)c.raw_hh() * sj, sj);
    T ld = 0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const MixView<T> mv = processed ? table_view<T>(c, K, j, tab) : regulate_to_slots<T>(c, K, j, p, sj, slots);
        T y, logd;
        gf_eval_logpdf<T>(mv, c.inv_type, x[j], y, logd);
        x[j] = y;
        ld += logd;
    }
    logdet += ld;
}

template <typename T, int DM>
JF_DEVINL void gf_layer_sample(T* x, T& logdet, const GfLayerC<T>& c, int d, int K, bool processed, const T* tab,
                               const T* p, int64_t sj, T* slots, int& n_evals, int& n_unconv) {
    T ld = 0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const MixView<T> mv = processed ? table_view<T>(c, K, j, tab) : regulate_to_slots<T>(c, K, j, p, sj, slots);
        T logd;
        int ev;
        bool conv;
        x[j] = gf_solve<T>(mv, c.inv_type, x[j], logd, ev, conv);
        ld += logd;
        n_evals += ev;
        n_unconv += conv ? 0 : 1;
    }
    logdet -= ld;
    if (c.hh_iter > 0)
        householder_apply<T, DM>(x, d, c.hh_iter, false, processed, tab + c.tab_hh(), p + (int64_t)c.raw_hh() * sj, sj);
    if (c.has_offset) {
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] += processed ? tab[c.tab_off + j] : p[(int64_t)(c.raw_off + j) * sj];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// "t": affine layer x = L z + offset with a lower-triangular L (reference layers/euclidean/multivariate_normal.py:228-272,
// layers/matrix_fns.py:4-145, offset handling euclidean_base.py:34-75).  Raw slice: [offset d][log-diagonal 1 | d][lower
// entries d(d-1)/2, stored sub-diagonal by sub-diagonal starting at the bottom-left corner].  The diagonal is the same
// bounded exponential as the "g" widths.  The reference inverts L through cofactors; forward substitution is the same
// map.  Parameters are read raw in both modes (p with stride sj; shared vectors have sj = 1).
// ---------------------------------------------------------------------------------------------------------------------
JF_DEVINL int mvn_lower_index(int d, int i, int j) {      // entry of L[i][j], i > j, inside `lower_triangular_entries`
    const int ind = d - 1 - (i - j);                        // which sub-diagonal (0 = bottom-left corner, 1 entry)
    return ind * (ind + 1) / 2 + j;
}

template <typename T, int DM>
JF_DEVINL void mvn_layer_logpdf(T* x, T& logdet, const GfLayerC<T>& c, int d, const T* p, int64_t sj) {
    if (c.has_offset) {
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] -= p[(int64_t)(c.raw_off + j) * sj];
    }
    const int cov = c.inv_type;
    if (cov == 0) return;
    const T* q = p + (int64_t)(c.raw_off + (c.has_offset ? d : 0)) * sj;
    if (cov == 1) {
        T w, iw;
        regulate_width(q[0], c.w_min, c.inv_w_max, w, iw);
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] *= iw;
        logdet -= T(d) * log(w);
        return;
    }
    T ld = 0;
#pragma unroll
    for (int i = 0; i < d; ++i) {
        T w, iw;
        regulate_width(q[(int64_t)i * sj], c.w_min, c.inv_w_max, w, iw);
        T acc = x[i];
        if (cov == 3) {
#pragma unroll
            for (int j = 0; j < i; ++j) acc = fma(-q[(int64_t)(d + mvn_lower_index(d, i, j)) * sj], x[j], acc);
        }
        x[i] = acc * iw;
        ld += log(w);
    }
    logdet -= ld;
}

template <typename T, int DM>
JF_DEVINL void mvn_layer_sample(T* x, T& logdet, const GfLayerC<T>& c, int d, const T* p, int64_t sj) {
    const int cov = c.inv_type;
    const T* q = p + (int64_t)(c.raw_off + (c.has_offset ? d : 0)) * sj;
    if (cov == 1) {
        T w, iw;
        regulate_width(q[0], c.w_min, c.inv_w_max, w, iw);
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] *= w;
        logdet += T(d) * log(w);
    } else if (cov >= 2) {
        T ld = 0;
#pragma unroll
        for (int ii = 0; ii < d; ++ii) {                    // bottom row first: x[i] only needs the old x[j], j <= i
            const int i = d - 1 - ii;
            T w, iw;
            regulate_width(q[(int64_t)i * sj], c.w_min, c.inv_w_max, w, iw);
            T acc = x[i] * w;
            if (cov == 3) {
#pragma unroll
                for (int j = 0; j < i; ++j) acc = fma(q[(int64_t)(d + mvn_lower_index(d, i, j)) * sj], x[j], acc);
            }
            x[i] = acc;
            ld += log(w);
        }
        logdet += ld;
    }
    if (c.has_offset) {
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] += p[(int64_t)(c.raw_off + j) * sj];
    }
}

// out-of-line wrappers for the chain kernels: "t" layers are rare next to "g" layers and their code (regulators, logs)
// would otherwise sit in the middle of the hot loop's instruction stream.  The row vector is copied so that the caller's
// registers do not become an address-taken local array.
template <typename T, int DM>
__device__ __noinline__ void mvn_layer_cold(bool logpdf, T* xs, T* logdet, const GfLayerC<T>* c, int d, const T* p, int64_t sj) {
    T x[DM];
#pragma unroll
    for (int j = 0; j < DM; ++j) x[j] = (j < d) ? xs[j] : T(0);
    T ld = *logdet;
    if (logpdf) mvn_layer_logpdf<T, DM>(x, ld, *c, d, p, sj);
    else mvn_layer_sample<T, DM>(x, ld, *c, d, p, sj);
#pragma unroll
    for (int j = 0; j < DM; ++j) if (j < d) xs[j] = x[j];
    *logdet = ld;
}

}  // namespace jf
