// Gaussianization-flow layer "g" with the NON-DEFAULT options of the reference (the option
// sweep of the reference's tests/test_general.py:307-320):
//   rotation_mode    none | angles (Givens chain) | cayley (d = 2) | triangular_combination   gaussianization_flow.py:711-800,
//                                                                                            :942-987, :1004-1055
//   width regulator  softplus_for_width, clamp_widths, width_smooth_saturation = 0, unbounded gaussianization_flow.py:264-317
//   center_mean      last mean from the others                                               gaussianization_flow.py:841-848
//   add_skewness     skewed logistic kernels sigma(a)^s, half of them mirrored               gaussianization_flow.py:363-386, :417-442,
//                                                                                            extra_functions.py:14-61
//   nonlinear_stretch_type = "rq_splines": rational-quadratic spline with linear tails       gaussianization_flow.py:863-909, :926-940,
//                                                                                            layers/spline_fns.py:188-358
// These run on their own chain kernel (gfx_chain_kernel) so that the specialised kernel of the default configuration
// (gf_chain_kernel, instruction-cache bound) is not touched.  Same execution model: one thread per row, the regulated
// parameters of the current (layer, dimension) in per-thread shared-memory slots (also for shared parameter vectors:
// every thread then reads the same raw values, a broadcast), register-resident root finding.  Layers with default
// options inside such a chain and "t" layers are handled here as well.
#pragma once
#include "subpdf_args.cuh"
#include "gf.cuh"
#include "spline.cuh"

namespace jf {

template <typename T>
struct GfxLayerC {
    GfLayerC<T> base;     // K, d, hh_iter, inv_type, norm_mode, has_offset, kind, raw_off, regulator bounds
    int rot_mode;         // JF_ROT_*
    int width_mode;       // JF_WIDTH_*
    int width_clamp;      // clamp the raw width parameter into [clamp_lo, clamp_hi] first
    int skew;             // add_skewness
    int center_mean;
    int stretch;          // JF_STRETCH_*
    // raw offsets (elements from the start of the sub-pdf parameter vector)
    //   classic: off_rot, off_m (means), off_w (log widths), off_n (log weights), off_s (log skew exponents)
    //   rq_splines: off_m = log_widths [d,K], off_w = log_heights [d,K], off_n = log_derivatives [d,K+1], off_s = boundary [d,4]
    int off_rot, off_m, off_w, off_n, off_s;
    int pad_;
    T clamp_lo, clamp_hi;
};

template <typename T>
struct GfxChainArgs {
    SubPdfArgs<T> a;
    GfxLayerC<T> layers[JF_MAX_LAYERS];
};

constexpr int kGfxFields = 5;   // slot fields per mixture kernel: mean, 1/width, weight (log weight when skewed), skew exponent, pdf prefactor

// ---------------------------------------------------------------------------------------------------------------------
// rotations
// ---------------------------------------------------------------------------------------------------------------------
// `inverse` = log_pdf direction (x <- Q^T x, or the inverse of the triangular product), else sampling (x <- Q x)
template <typename T>
__device__ __noinline__ void gfx_rotate(T* x, const GfxLayerC<T>& c, bool inverse, const T* p, int64_t sj) {
    const int d = c.base.d;
    const T* q = p + (int64_t)c.off_rot * sj;
    if (c.rot_mode == JF_ROT_HOUSEHOLDER) {
        if (c.base.hh_iter > 0) householder_apply<T, JF_MAX_DIM>(x, d, c.base.hh_iter, inverse, false, nullptr, q, sj);
        return;
    }
    if (c.rot_mode == JF_ROT_NONE || d < 2) return;
    if (c.rot_mode == JF_ROT_ANGLES) {
        // Q = G_{n-1} ... G_0 over the pairs (a,b) in itertools.combinations order; G[a,a]=G[b,b]=cos, G[a,b]=sin, G[b,a]=-sin
        const int n = d * (d - 1) / 2;
        if (!inverse) {
            int idx = 0;
            for (int a = 0; a < d - 1; ++a)
                for (int b = a + 1; b < d; ++b, ++idx) {
                    T s, co;
                    sincos(q[(int64_t)idx * sj], &s, &co);
                    const T xa = x[a], xb = x[b];
                    x[a] = co * xa + s * xb;
                    x[b] = co * xb - s * xa;
                }
        } else {
            int idx = n - 1;
            for (int a = d - 2; a >= 0; --a)
                for (int b = d - 1; b > a; --b, --idx) {
                    T s, co;
                    sincos(q[(int64_t)idx * sj], &s, &co);
                    const T xa = x[a], xb = x[b];
                    x[a] = co * xa - s * xb;
                    x[b] = co * xb + s * xa;
                }
        }
        return;
    }
    if (c.rot_mode == JF_ROT_CAYLEY) {                          // d == 2: [[c, -s], [s, c]]
        const T r = q[0];
        const T f = T(1) / (T(1) + r * r);
        const T co = (T(1) - r * r) * f, s = T(2) * r * f;
        const T x0 = x[0], x1 = x[1];
        if (!inverse) { x[0] = co * x0 - s * x1; x[1] = s * x0 + co * x1; }
        else          { x[0] = co * x0 + s * x1; x[1] = co * x1 - s * x0; }
        return;
    }
    // triangular_combination: x <- L (exp(diag) * (U x)); L unit lower from the first d(d-1)/2 entries, diag from the
    // next d-1 (the last one makes the sum zero: volume preserving), U = transpose of the unit lower matrix of the rest
    const int npm = d * (d - 1) / 2;
    const T* ql = q;
    const T* qm = q + (int64_t)npm * sj;
    const T* qr = q + (int64_t)(npm + d - 1) * sj;
    T dsum = 0;
    if (!inverse) {
        for (int j = 0; j < d; ++j) {                            // U x, ascending (x[i], i > j, still old)
            T acc = x[j];
            for (int i = j + 1; i < d; ++i) acc = fma(qr[(int64_t)mvn_lower_index(d, i, j) * sj], x[i], acc);
            x[j] = acc;
        }
        for (int j = 0; j < d; ++j) {
            const T dg = (j < d - 1) ? qm[(int64_t)j * sj] : -dsum;
            dsum += dg;
            x[j] *= exp(dg);
        }
        for (int i = d - 1; i >= 0; --i) {                       // L x, descending
            T acc = x[i];
            for (int j = 0; j < i; ++j) acc = fma(ql[(int64_t)mvn_lower_index(d, i, j) * sj], x[j], acc);
            x[i] = acc;
        }
    } else {
        for (int i = 0; i < d; ++i) {                            // L^-1: forward substitution
            T acc = x[i];
            for (int j = 0; j < i; ++j) acc = fma(-ql[(int64_t)mvn_lower_index(d, i, j) * sj], x[j], acc);
            x[i] = acc;
        }
        for (int j = 0; j < d; ++j) {
            const T dg = (j < d - 1) ? qm[(int64_t)j * sj] : -dsum;
            dsum += dg;
            x[j] /= exp(dg);
        }
        for (int j = d - 1; j >= 0; --j) {                       // U^-1: back substitution
            T acc = x[j];
            for (int i = j + 1; i < d; ++i) acc = fma(-qr[(int64_t)mvn_lower_index(d, i, j) * sj], x[i], acc);
            x[j] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// parameter regulation into the per-thread slots
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
JF_DEVINL T gfx_inv_width(T raw, const GfxLayerC<T>& c) {
    if (c.width_clamp) raw = tmin(tmax(raw, c.clamp_lo), c.clamp_hi);
    if (c.width_mode == JF_WIDTH_SMOOTH) return regulate_inv_width(raw, c.base.w_min, c.base.inv_w_max);
    const T w = (c.width_mode == JF_WIDTH_SOFTPLUS ? softplus_t(raw) : exp(raw)) + c.base.w_min;
    return T(1) / w;
}

template <typename T>
struct GfxView {
    MixView<T> mv;
    unsigned s, pre;      // byte addresses of the skew exponents / pdf prefactors (log(s/w) + log weight), k = 0
};

template <typename T>
__device__ __noinline__ GfxView<T> gfx_regulate(const GfxLayerC<T>& c, int j, const T* p, int64_t sj, T* slots) {
    const int d = c.base.d, K = c.base.K, nt = blockDim.x;
    const size_t fs = (size_t)K * nt;
    T* sm = slots + threadIdx.x;
    T* si = sm + fs;
    T* sn = si + fs;
    T* ss = sn + fs;
    T* sp = ss + fs;
    const int Km = K - (c.center_mean ? 1 : 0);
    for (int k = 0; k < K; ++k) {
        if (k < Km) sm[(size_t)k * nt] = p[(int64_t)(c.off_m + k * d + j) * sj];
        si[(size_t)k * nt] = gfx_inv_width(p[(int64_t)(c.off_w + k * d + j) * sj], c);
        sn[(size_t)k * nt] = (c.base.norm_mode != JF_NORM_NONE) ? p[(int64_t)(c.off_n + k * d + j) * sj] : T(0);
    }
    T nmax = -Num<T>::big;
    if (c.base.norm_mode == JF_NORM_RAW)
        for (int k = 0; k < K; ++k) nmax = tmax(nmax, sn[(size_t)k * nt]);
    T nsum = 0;
    for (int k = 0; k < K; ++k) {
        T g;
        if (c.base.norm_mode == JF_NORM_REGULATED) g = regulate_norm(sn[(size_t)k * nt], c.base.n_min, c.base.n_max);
        else if (c.base.norm_mode == JF_NORM_RAW) g = exp(sn[(size_t)k * nt] - nmax);
        else g = T(1);
        nsum += g;
        sn[(size_t)k * nt] = g;
    }
    if (c.center_mean) {
        // last mean = -(sum_{k<K-1} m_k n_k)/n_{K-1} with the (unnormalised) linear weights: the ratio is scale free
        T acc = 0;
        for (int k = 0; k < K - 1; ++k) acc = fma(sm[(size_t)k * nt], sn[(size_t)k * nt], acc);
        sm[(size_t)(K - 1) * nt] = -acc / sn[(size_t)(K - 1) * nt];
    }
    T mmin = Num<T>::big, mmax = -Num<T>::big;
    const T inv = T(1) / nsum;
    for (int k = 0; k < K; ++k) {
        const T m = sm[(size_t)k * nt];
        mmin = tmin(mmin, m);
        mmax = tmax(mmax, m);
        const T n = sn[(size_t)k * nt] * inv;
        if (c.skew) {
            // exponent regulated into [0.1, 9.1] like the widths (gaussianization_flow.py:384)
            const T e = T(0.1) + T(1) / (T(1.0 / 9.0) + exp(-p[(int64_t)(c.off_s + k * d + j) * sj]));
            const T ln = log(n);
            ss[(size_t)k * nt] = e;
            sp[(size_t)k * nt] = log(si[(size_t)k * nt]) + log(e) + ln;
            sn[(size_t)k * nt] = ln;
        } else {
            sn[(size_t)k * nt] = n;
        }
    }
    GfxView<T> v;
    v.mv.m = smem_addr(sm); v.mv.iw = smem_addr(si); v.mv.n = smem_addr(sn);
    v.mv.skb = (unsigned)(nt * sizeof(T)); v.mv.K = K; v.mv.mmin = mmin; v.mv.mmax = mmax;
    v.s = smem_addr(ss); v.pre = smem_addr(sp);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// skewed mixture in log space, term by term as the reference (gaussianization_flow.py:389-454): the branch thresholds
// of softplus and of log(((1+e^x)^a - 1)/(1+e^x)^a) are part of the numerical contract
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
struct Lse {              // online log-sum-exp
    T m, s;
    JF_DEVINL void init() { m = -Num<T>::big; s = 0; }
    JF_DEVINL void add(T t) {
        if (t > m) { s = s * exp(m - t) + T(1); m = t; }
        else s += exp(t - m);
    }
    JF_DEVINL T value() const { return m + log(s); }
};

// `sfx` = 1 - cdf_ref accumulated term by term without cancellation, where cdf_ref is the cdf exactly as the reference
// forms it (branch approximations included).  The reference's inverse-normal stage is a function of its cdf alone
// (sqrt2*erfinv(2cdf-1)); its own log_sf differs from log(1-cdf) by up to e^-20 because of those branches, so the
// upper-tail evaluation (erfcinv) must not use it.
template <typename T>
__device__ __noinline__ void skew_eval(const GfxView<T>& v, T x, T& lc, T& ls, T& lp, T& sfx) {
    const int K = v.mv.K;
    Lse<T> C, S, P;
    C.init(); S.init(); P.init();
    T sf = 0;
    for (int k = 0; k < K; ++k) {
        const T a = (x - mv_m(v.mv, k)) * mv_iw(v.mv, k);
        const bool pos = k < K / 2;
        const T b = pos ? a : -a;
        const T ln = mv_n(v.mv, k);
        const T e = lds(v.s + k * v.mv.skb, T());
        const T spn = softplus_t(-b);
        const T sp = e * spn;
        P.add(-b + lds(v.pre + k * v.mv.skb, T()) - (e + T(1)) * spn);
        // complement of sigma(b)^e, and 1 - (that complement) for the mirrored kernels
        T r, one_minus_cmp;
        if (-b <= T(-20)) { r = log(e) - b; one_minus_cmp = T(1) - exp(r - sp); }
        else if (sp > T(20)) { r = sp; one_minus_cmp = T(0); }
        else if (sp < T(1e-8)) { r = log(sp); one_minus_cmp = T(1) - sp * exp(-sp); }
        else { r = log(exp(sp) - T(1)); one_minus_cmp = exp(-sp); }
        const T t_pow = -sp + ln, t_cmp = (r - sp) + ln;
        C.add(pos ? t_pow : t_cmp);
        S.add(pos ? t_cmp : t_pow);
        sf = fma(exp(ln), pos ? -expm1(-sp) : one_minus_cmp, sf);
    }
    lc = C.value(); ls = S.value(); lp = P.value();
    sfx = sf;
}

// inverse-CDF stage from log cdf / log sf / log pdf (reference gaussianization_flow.py:480-671); sfx: see skew_eval
template <typename T>
JF_DEVINL void inv_stage_log(int type, T lc, T ls, T lp, T sfx, T& y, T& logd) {
    if (type == JF_INV_ISIGMOID) {
        y = lc - ls;
        const T hi = tmax(-ls, -lc);
        logd = hi + log1p(exp(-fabs(ls - lc))) + lp;
        return;
    }
    inv_stage_inormal<T>(type, lc, ls, lp, sfx, y, logd);
}

// x with sigma(x) = exp(u), u < 0
// x with sigma(x)^e = P, given lprob = log P (<= 0) and log_neg_lprob = log(-log P) (needed when P is so close to 1
// that -log P underflows next to 1: then 1 - e^u ~ -u and x = -log(-u))
template <typename T>
JF_DEVINL T quantile_logit(T lprob, T log_neg_lprob, T e) {
    const T u = lprob / e;
    if (-u < T(1e-8)) return -(log_neg_lprob - log(e));
    return u - log(-expm1(u));
}

// Root of y(x) = z for a skewed mixture: analytic bracket from the per-kernel quantiles (the mixture cdf is a convex
// combination of the kernel cdfs), then safeguarded Newton on y itself with bisection fall-back.
template <typename T>
__device__ __noinline__ T skew_solve(const GfxView<T>& v, int type, T z, T& logd_out, int& evals, bool& converged) {
    const int K = v.mv.K;
    T t_lo = z, t_hi = z;
    if (type != JF_INV_ISIGMOID) {
        const T marg = (type == JF_INV_PARTLY_CRUDE ? T(0.6) : T(0.05)) + T(0.02) * fabs(z);
        t_lo = logit_phi<T>(z - marg);
        t_hi = logit_phi<T>(z + marg);
    }
    T lo = Num<T>::big, hi = -Num<T>::big, x = 0, wmax = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const T t = pass == 0 ? t_lo : t_hi;
        // log p = log sigma(t) = -softplus(-t) and log(1-p) = -softplus(t), each without cancellation, plus the logs of
        // their magnitudes (|t| reaches 1e3 behind a non-orthogonal "rotation": 1-p or p is then ~e^-|t|)
        const T at = fabs(t);
        const T l1 = log1p(exp(-at));
        const T tiny_l = (at > T(30)) ? -at : log(l1);          // log of softplus(-|t|)
        const T lp_ = t > T(0) ? -l1 : (t - l1), llp = t > T(0) ? tiny_l : log(l1 - t);
        const T lq_ = t > T(0) ? (-t - l1) : -l1, llq = t > T(0) ? log(t + l1) : tiny_l;
        for (int k = 0; k < K; ++k) {
            const T m = mv_m(v.mv, k), w = T(1) / mv_iw(v.mv, k);
            const T e = lds(v.s + k * v.mv.skb, T());
            const bool pos = k < K / 2;
            const T a = pos ? quantile_logit(lp_, llp, e) : -quantile_logit(lq_, llq, e);
            const T c = fma(a, w, m);
            if (pass == 0) lo = tmin(lo, c); else hi = tmax(hi, c);
            x = fma(T(0.5) * exp(mv_n(v.mv, k)), c, x);
            wmax = tmax(wmax, w);
        }
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
#pragma unroll 1
    for (int it = 0; it < 100; ++it) {
        T lc, ls, lp, sfx, y;
        skew_eval(v, x, lc, ls, lp, sfx);
        ++evals;
        inv_stage_log(type, lc, ls, lp, sfx, y, logd);
        f = y - z;
        if (f < T(0)) lo = x; else hi = x;
        const T dx = f / exp(logd);
        const T xn = x - dx;
        const bool inside = (xn > lo) && (xn < hi);
        if (fabs(dx) <= tol_abs + tol_rel * fabs(x)) {
            // a step this small changes log y' by less than the tolerance: no confirming evaluation
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

// ---------------------------------------------------------------------------------------------------------------------
// nonlinear_stretch_type = "rq_splines": one rational-quadratic spline per dimension between [left,right] -> [bottom,top]
// with linear tails (layers/spline_fns.py:188-358; boundaries gaussianization_flow.py:895-907)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __noinline__ void gfx_rqs(const GfxLayerC<T>& c, int j, const T* p, int64_t sj, bool inverse, T x, T& out, T& lad) {
    const int K = c.base.K;
    T kx[JF_MAX_KDE + 1], ky[JF_MAX_KDE + 1];
    const T* pb = p + (int64_t)(c.off_s + j * 4) * sj;
    const T left = pb[0], right = left + exp(pb[sj]) + T(0.5);
    const T bottom = pb[2 * sj], top = bottom + exp(pb[3 * sj]) + T(0.5);
    for (int which = 0; which < 2; ++which) {
        const T* q = p + (int64_t)((which == 0 ? c.off_m : c.off_w) + j * K) * sj;
        T* kn = which == 0 ? kx : ky;
        const T lo = which == 0 ? left : bottom, hi = which == 0 ? right : top;
        T m = q[0];
        for (int k = 1; k < K; ++k) m = tmax(m, q[(int64_t)k * sj]);
        T sum = 0;
        for (int k = 0; k < K; ++k) { kn[k + 1] = exp(q[(int64_t)k * sj] - m); sum += kn[k + 1]; }
        T cum = 0;
        kn[0] = lo;                                           // (hi - lo) * 0 + lo
        for (int k = 0; k < K; ++k) {
            cum += T(1e-3) + (T(1) - T(1e-3) * T(K)) * (kn[k + 1] / sum);
            kn[k + 1] = (hi - lo) * cum + lo;
        }
    }
    const T* qd = p + (int64_t)(c.off_n + j * (K + 1)) * sj;
    const T* ks = inverse ? ky : kx;
    int idx = -1;
    for (int i = 0; i <= K; ++i) idx += (x >= ks[i]) ? 1 : 0;
    idx = idx < 0 ? 0 : (idx > K - 1 ? K - 1 : idx);
    const T dfirst = T(1e-3) + softplus_t(qd[0]), dlast = T(1e-3) + softplus_t(qd[(int64_t)K * sj]);
    const T d0 = T(1e-3) + softplus_t(qd[(int64_t)idx * sj]), d1 = T(1e-3) + softplus_t(qd[(int64_t)(idx + 1) * sj]);
    const T x0 = kx[idx], wk = kx[idx + 1] - kx[idx];
    const T y0 = ky[idx], hk = ky[idx + 1] - ky[idx];
    const T s = hk / wk;
    const T t = d0 + d1 - T(2) * s;
    T xi;
    if (inverse) {
        const T dy = x - y0;
        const T qa = dy * t + hk * (s - d0);
        const T qb = hk * d0 - dy * t;
        const T qc = -s * dy;
        xi = (T(2) * qc) / (-qb - sqrt(qb * qb - T(4) * qa * qc));
        out = xi * wk + x0;
    } else {
        xi = (x - x0) / wk;
    }
    const T xx = xi * (T(1) - xi);
    const T den = s + t * xx;
    if (!inverse) out = y0 + hk * (s * xi * xi + d0 * xx) / den;
    const T num = s * s * (d1 * xi * xi + T(2) * s * xx + d0 * (T(1) - xi) * (T(1) - xi));
    const T l = log(num) - T(2) * log(den);
    lad = inverse ? -l : l;
    if (!inverse) {
        if (x <= left)  { out = x * dfirst + (ky[0] - kx[0] * dfirst); lad = log(dfirst); }
        if (x >= right) { out = x * dlast + (ky[K] - kx[K] * dlast);   lad = log(dlast); }
    } else {
        if (x <= bottom) { out = x / dfirst + (kx[0] - ky[0] / dfirst); lad = -log(dfirst); }
        if (x >= top)    { out = x / dlast + (kx[K] - ky[K] / dlast);   lad = -log(dlast); }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// one layer on one row
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
JF_DEVINL void gfx_layer_logpdf(T* x, T& logdet, const GfxLayerC<T>& c, const T* p, int64_t sj, T* slots) {
    const int d = c.base.d;
    if (c.base.has_offset)
        for (int j = 0; j < d; ++j) x[j] -= p[(int64_t)(c.base.raw_off + j) * sj];
    gfx_rotate<T>(x, c, true, p, sj);
    T ld = 0;
    for (int j = 0; j < d; ++j) {
        T y, logd;
        if (c.stretch == JF_STRETCH_RQS) {
            gfx_rqs<T>(c, j, p, sj, false, x[j], y, logd);
        } else {
            const GfxView<T> v = gfx_regulate<T>(c, j, p, sj, slots);
            if (c.skew) {
                T lc, ls, lp, sfx;
                skew_eval<T>(v, x[j], lc, ls, lp, sfx);
                inv_stage_log<T>(c.base.inv_type, lc, ls, lp, sfx, y, logd);
            } else {
                gf_eval_logpdf<T>(v.mv, c.base.inv_type, x[j], y, logd);
            }
        }
        x[j] = y;
        ld += logd;
    }
    logdet += ld;
}

template <typename T>
JF_DEVINL void gfx_layer_sample(T* x, T& logdet, const GfxLayerC<T>& c, const T* p, int64_t sj, T* slots, int& n_evals,
                                int& n_unconv) {
    const int d = c.base.d;
    T ld = 0;
    for (int j = 0; j < d; ++j) {
        T logd;
        if (c.stretch == JF_STRETCH_RQS) {
            T y;
            gfx_rqs<T>(c, j, p, sj, true, x[j], y, logd);     // returns minus the forward log-derivative
            x[j] = y;
            ld -= logd;
        } else {
            const GfxView<T> v = gfx_regulate<T>(c, j, p, sj, slots);
            int ev;
            bool conv;
            x[j] = c.skew ? skew_solve<T>(v, c.base.inv_type, x[j], logd, ev, conv)
                          : gf_solve<T>(v.mv, c.base.inv_type, x[j], logd, ev, conv);
            ld += logd;
            n_evals += ev;
            n_unconv += conv ? 0 : 1;
        }
    }
    logdet -= ld;
    gfx_rotate<T>(x, c, false, p, sj);
    if (c.base.has_offset)
        for (int j = 0; j < d; ++j) x[j] += p[(int64_t)(c.base.raw_off + j) * sj];
}

// Euclidean sub-pdf with at least one non-default "g" layer.  Dynamic shared memory: kGfxFields*Kmax*blockDim.x slots.
template <typename T, int DIR>
__global__ void __launch_bounds__(256) gfx_chain_kernel(const __grid_constant__ GfxChainArgs<T> g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* slots = reinterpret_cast<T*>(smem_raw);
    const SubPdfArgs<T>& a = g.a;
    const int d = a.d;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.B) return;
    T x[JF_MAX_DIM];
    for (int j = 0; j < d; ++j) x[j] = a.in[row * a.ld_in + j];
    T logdet = a.logdet_in ? a.logdet_in[row] : T(0);
    const T* prow = a.params + row * a.sr;
    T zsq = 0;
    if (DIR == JF_DIR_LOGPDF) {
        if (a.emb_out)
            for (int j = 0; j < d; ++j) a.emb_out[row * a.ld_emb + j] = x[j];
        for (int l = a.n_layers - 1; l >= 0; --l) {
            if (g.layers[l].base.kind == 1) mvn_layer_logpdf<T, JF_MAX_DIM>(x, logdet, g.layers[l].base, d, prow, a.sj);
            else gfx_layer_logpdf<T>(x, logdet, g.layers[l], prow, a.sj, slots);
        }
        for (int j = 0; j < d; ++j) zsq = fma(x[j], x[j], zsq);
    } else {
        for (int j = 0; j < d; ++j) zsq = fma(x[j], x[j], zsq);
        int n_evals = 0, n_unconv = 0;
        for (int l = 0; l < a.n_layers; ++l) {
            if (g.layers[l].base.kind == 1) mvn_layer_sample<T, JF_MAX_DIM>(x, logdet, g.layers[l].base, d, prow, a.sj);
            else gfx_layer_sample<T>(x, logdet, g.layers[l], prow, a.sj, slots, n_evals, n_unconv);
        }
        if (a.emb_out)
            for (int j = 0; j < d; ++j) a.emb_out[row * a.ld_emb + j] = x[j];
        if (n_unconv) status_add(a.status, JF_STATUS_UNCONVERGED, n_unconv);
        status_add_warp(a.status, JF_STATUS_ITERATIONS, n_evals);
    }
    bool bad = !finite_(logdet);
    for (int j = 0; j < d; ++j) {
        a.out[row * a.ld_out + j] = x[j];
        bad = bad || !finite_(x[j]);
    }
    if (bad) status_add(a.status, JF_STATUS_NONFINITE, 1);
    if (a.logdet_out) a.logdet_out[row] = logdet;
    if (a.logbase_out) {
        const T prev = a.logbase_in ? a.logbase_in[row] : T(0);
        a.logbase_out[row] = prev - T(0.5) * zsq - T(d) * T(kLogSqrt2Pi);
    }
}

// defined in gfx_inst.cu
template <typename T>
int launch_gfx(const GfxChainArgs<T>& g, int direction, int kmax, cudaStream_t st);

}  // namespace jf
