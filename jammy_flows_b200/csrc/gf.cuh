// Gaussianization-flow layer "g": device math.
//
// What it computes is fixed by the reference (layers/euclidean/gaussianization_flow.py, cited per function); HOW is
// B200-first: one thread owns one row, the K mixture parameters of the current dimension live in registers, the
// logistic mixture is evaluated in LINEAR space with a single rescaling exponent (1 exp + 1 reciprocal per kernel
// instead of the reference's softplus + 3 logsumexp = 4 exp + 1 log1p per kernel), and the sampling direction is a
// register-resident bracketed Newton iteration (no [B,K,d] temporaries, no host sync per iteration).
#pragma once
#include "common.cuh"

namespace jf {

// ---------------------------------------------------------------------------------------------------------------------
// Layer constants (host fills from JfLayerDesc)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
struct GfLayerC {
    int K, d, hh_iter, inv_type, norm_mode, has_offset;
    int raw_off;   // start of the layer's slice in the raw parameter vector
    int tab_off;   // start of the layer's block in the processed shared-memory table
    T w_min, inv_w_max, n_min, n_max;
    // raw slice: [offset d][vs hh_iter*d][means K*d][log_w K*d][log_n K*d]   (index inside K*d blocks: k*d + j)
    __host__ __device__ int raw_hh() const { return raw_off + (has_offset ? d : 0); }
    __host__ __device__ int raw_m() const { return raw_hh() + hh_iter * d; }
    __host__ __device__ int raw_w() const { return raw_m() + K * d; }
    __host__ __device__ int raw_n() const { return raw_w() + K * d; }
    // processed table block: [offset d][vhat hh_iter*d][m K*d][w K*d][iw K*d][n K*d]
    __host__ __device__ int tab_hh() const { return tab_off + d; }
    __host__ __device__ int tab_m() const { return tab_hh() + hh_iter * d; }
    __host__ __device__ int tab_w() const { return tab_m() + K * d; }
    __host__ __device__ int tab_iw() const { return tab_w() + K * d; }
    __host__ __device__ int tab_n() const { return tab_iw() + K * d; }
    __host__ __device__ int tab_size() const { return d + hh_iter * d + 4 * K * d; }
};

// ---------------------------------------------------------------------------------------------------------------------
// Parameter regulation (reference gaussianization_flow.py:23-47, 300-317, 342, 406)
// ---------------------------------------------------------------------------------------------------------------------
// width:  log_w = log(w_min + 1/(1/w_max + exp(-raw)))  ->  w and 1/w
template <typename T>
JF_DEVINL void regulate_width(T raw, T w_min, T inv_w_max, T& w, T& iw) {
    T q = inv_w_max + exp(-raw);
    w = w_min + T(1) / q;
    iw = T(1) / w;
}
// norm (unnormalised, linear space): exp(log_n_regulated) = n_min + n_max * sigmoid(raw)
template <typename T>
JF_DEVINL T regulate_norm(T raw, T n_min, T n_max) {
    return n_min + n_max / (T(1) + exp(-raw));
}

template <typename T, int KM>
struct Mix {
    T m[KM], w[KM], iw[KM], n[KM];
};

// Load the K mixture parameters of dimension j for this thread's row.
//   processed: from the shared-memory table (already regulated; stride 1)
//   raw:       from global/shared raw parameters (element i at p[i*sj]) and regulate on the fly
template <typename T, int KM>
JF_DEVINL void load_mix(Mix<T, KM>& mx, const GfLayerC<T>& c, int K, int j, bool processed, const T* tab,
                        const T* p, int64_t sj) {
    const int d = c.d;
    if (processed) {
        const T* tm = tab + c.tab_m() + j;
        const T* tw = tab + c.tab_w() + j;
        const T* ti = tab + c.tab_iw() + j;
        const T* tn = tab + c.tab_n() + j;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            mx.m[k] = tm[k * d];
            mx.w[k] = tw[k * d];
            mx.iw[k] = ti[k * d];
            mx.n[k] = tn[k * d];
        }
    } else {
        const T* pm = p + (int64_t)(c.raw_m() + j) * sj;
        const T* pw = p + (int64_t)(c.raw_w() + j) * sj;
        const T* pn = p + (int64_t)(c.raw_n() + j) * sj;
        T rw[KM], rn[KM];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            mx.m[k] = pm[(int64_t)k * d * sj];
            rw[k] = pw[(int64_t)k * d * sj];
            rn[k] = (c.norm_mode != JF_NORM_NONE) ? pn[(int64_t)k * d * sj] : T(0);
        }
        T nsum = 0, nmax = -Num<T>::big;
        if (c.norm_mode == JF_NORM_RAW) {
#pragma unroll
            for (int k = 0; k < K; ++k) nmax = tmax(nmax, rn[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            regulate_width(rw[k], c.w_min, c.inv_w_max, mx.w[k], mx.iw[k]);
            T g;
            if (c.norm_mode == JF_NORM_REGULATED) g = regulate_norm(rn[k], c.n_min, c.n_max);
            else if (c.norm_mode == JF_NORM_RAW) g = exp(rn[k] - nmax);
            else g = T(1);
            mx.n[k] = g;
            nsum += g;
        }
        T inv = T(1) / nsum;
#pragma unroll
        for (int k = 0; k < K; ++k) mx.n[k] *= inv;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K-logistic mixture: log CDF / log SF / log PDF  (reference gaussianization_flow.py:389-454)
//   a_k = (x - m_k)/w_k ;  cdf = sum n_k sigma(a_k) ; sf = sum n_k sigma(-a_k) ; pdf = sum n_k sigma(a_k) sigma(-a_k)/w_k
// Linear-space evaluation with ONE shared rescaling exponent delta = min_k |a_k| when all a_k have the same sign
// (x outside the hull of the means), so nothing underflows however far out x is:
//   u_k = exp(delta - |a_k|) in (0,1],  e_k = exp(-|a_k|) = u_k * exp(-delta),  r_k = 1/(1+e_k)
//   sigma(|a|) = r,  sigma(-|a|) = e r.   Results are returned as (S, shift) with log X = log S - shift.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
struct MixVal {
    T Sc, Ss, Sp;   // rescaled cdf / sf / pdf sums (Ss is the EXACT survival sum)
    T dc, ds, dp;   // shifts: log cdf = log Sc - dc, ...
    T ex;           // softplus-threshold excess of the reference's sf: sf_ref = Ss + ex  (see mix_eval)
};

// Reference quirk that is part of the numerical contract: F.softplus(t) returns t for t > 20 (torch default
// threshold), so for kernels with a_k < -20 the reference's log-terms drop the factor 1/(1+e^{a_k}):
//   cdf_k = n e^{a}        (exact n e^{a} r)      -- a 2e-9 relative change of a term that is itself <= 2e-9: invisible
//   sf_k  = n              (exact n r)            -- sf_ref = sf_exact + ex,  ex = sum_{a_k<-20} n_k e^{a_k} r_k <= 2e-9
//   pdf_k = n e^{a}/w      (exact n e^{a} r^2/w)
// so that cdf_ref + sf_ref = 1 + ex.  The kernel carries the exact sf (needed for an accurate Phi^-1 from the upper
// tail) and the excess separately.
template <typename T, int KM>
JF_DEVINL MixVal<T> mix_eval(const Mix<T, KM>& mx, int K, T x) {
    T a[KM];
    T amax = -Num<T>::big, amin = Num<T>::big;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        a[k] = (x - mx.m[k]) * mx.iw[k];
        amax = tmax(amax, a[k]);
        amin = tmin(amin, a[k]);
    }
    const bool all_neg = amax < T(0), all_pos = amin > T(0);
    const T delta = all_neg ? -amax : (all_pos ? amin : T(0));
    const T E = exp(-delta);
    T Sc = 0, Ss = 0, Sp = 0, ex = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const T u = exp(delta - fabs(a[k]));
        const T e = u * E;
        const T rx = T(1) / (T(1) + e);            // exact sigma(|a|)
        const bool quirk = a[k] < T(-20);
        const T r = quirk ? T(1) : rx;
        const T ur = u * r;
        const bool pos = a[k] >= T(0);
        Sc = fma(mx.n[k], pos ? r : ur, Sc);
        Ss = fma(mx.n[k], pos ? ur : rx, Ss);
        Sp = fma(mx.n[k] * mx.iw[k], ur * r, Sp);
        if (quirk) ex = fma(mx.n[k], e * rx, ex);
    }
    MixVal<T> v;
    v.Sc = Sc; v.Ss = Ss; v.Sp = Sp; v.ex = ex;
    v.dc = all_neg ? delta : T(0);
    v.ds = all_pos ? delta : T(0);
    v.dp = delta;
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// inverse-CDF stage and its log-derivative (reference gaussianization_flow.py:480-560 and :568-671)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
JF_DEVINL void inv_stage(int type, const MixVal<T>& v, T& y, T& logd) {
    const T lc = log(v.Sc) - v.dc, ls = log(v.Ss + v.ex) - v.ds, lp = log(v.Sp) - v.dp;
    if (type == JF_INV_ISIGMOID) {
        y = lc - ls;
        // logaddexp(-ls,-lc) + lp = lp - lc - ls + log(cdf+sf), and cdf_ref + sf_ref = 1 + ex (ex <= 2e-9)
        logd = lp - lc - ls + v.ex;
        return;
    }
    const T eps = T(0.5e-7), pa = T(0.147);
    const T pc = T(2.0 / (kPi * 0.147));
    const T cdf = exp(lc);
    const T lnf = lc + ls + T(1.3862943611198906);   // log 4
    const T F = pc + lnf * T(0.5);
    const T F2 = sqrt(F * F - lnf / pa);
    // upper bulk limit tested on the exact survival function (cdf < 1-eps <=> sf > eps): in fp32 1-eps rounds to 1 and
    // a cdf that rounds to 1-ulp would otherwise enter the bulk branch with an underflowed tail
    const T sfx = v.Ss * exp(-v.ds);
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

// value and plain derivative for the Newton iteration (reference :685-695)
template <typename T>
JF_DEVINL void inv_stage_newton(int type, const MixVal<T>& v, T& y, T& dy) {
    if (type == JF_INV_ISIGMOID) {
        const T ssq = v.Ss + v.ex;
        y = log(v.Sc / ssq) + (v.ds - v.dc);
        dy = v.Sp / (v.Sc * ssq);   // exp(lp - lc - ls): the shifts cancel exactly (dp = dc + ds)
        return;
    }
    T logd;
    inv_stage(type, v, y, logd);
    dy = exp(logd);
}

// log(Phi(t)/Phi(-t)): the logistic-scale target that corresponds to a Gaussian-scale target t
template <typename T>
JF_DEVINL T logit_phi(T t) {
    const T at = fabs(t);
    const T lim = sizeof(T) == 8 ? T(25) : T(12);
    if (at < lim) {
        const T r = T(0.70710678118654752);
        return log(erfc(-t * r)) - log(erfc(t * r));
    }
    const T v = T(0.5) * at * at + log(at) + T(kLogSqrt2Pi);
    return t > 0 ? v : -v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Root of  y(x) = z  for one element (reference: 25 bisections on [-1e5,1e5] + <=20 masked Newton steps,
// gaussianization_flow.py:921 + bisection_n_newton.py:11-135).  Same root, different trajectory:
//   * analytic bracket: with t the logistic-scale target, every kernel satisfies sigma((x-m_k)/w_k) <= sigma(t) for
//     x <= min_k(m_k + t w_k) and >= sigma(t) for x >= max_k(m_k + t w_k), so the mixture root lies in between;
//   * safeguarded Newton inside the bracket (bisect when a step leaves it or |f| does not shrink);
//   * stops when |dx| <= 1e-14 + 4 eps |x| (fp64), i.e. at least as tight as the reference's 1e-14 row-sum test.
// Returns the root, the log-derivative at the root and the number of function evaluations.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int KM>
JF_DEVINL T gf_solve(const Mix<T, KM>& mx, int K, int type, T z, T& logd_out, int& evals, bool& converged) {
    // logistic-scale targets with safety margins for the non-exact inverse-CDF variants
    T t_lo, t_hi;
    if (type == JF_INV_ISIGMOID) {
        t_lo = t_hi = z;
    } else {
        const T marg = (type == JF_INV_PARTLY_CRUDE ? T(0.6) : T(0.05)) + T(0.02) * fabs(z);
        t_lo = logit_phi(z - marg);
        t_hi = logit_phi(z + marg);
    }
    T lo = Num<T>::big, hi = -Num<T>::big, x = 0, wmax = 0;
    const T t_mid = T(0.5) * (t_lo + t_hi);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        lo = tmin(lo, fma(t_lo, mx.w[k], mx.m[k]));
        hi = tmax(hi, fma(t_hi, mx.w[k], mx.m[k]));
        x = fma(mx.n[k], fma(t_mid, mx.w[k], mx.m[k]), x);
        wmax = tmax(wmax, mx.w[k]);
    }
    {
        const T pad = T(1e-3) * wmax + T(64) * Num<T>::eps * (fabs(lo) + fabs(hi));
        lo -= pad;
        hi += pad;
        // keep inside the reference's search interval
        lo = tmax(lo, T(-1e5));
        hi = tmin(hi, T(1e5));
        x = clampv(x, lo, hi);
    }
    const T tol_abs = Num<T>::newton_abs_tol, tol_rel = T(4) * Num<T>::eps;
    T fprev = Num<T>::big;
    T logd = 0, f = 0;
    converged = false;
    evals = 0;
    MixVal<T> v;
    const int kMaxIt = 64;
    for (int it = 0; it < kMaxIt; ++it) {
        v = mix_eval<T, KM>(mx, K, x);
        ++evals;
        T y, dy;
        inv_stage_newton(type, v, y, dy);
        f = y - z;
        if (f < T(0)) lo = x; else hi = x;
        const T dx = f / dy;
        T xn = x - dx;
        const bool inside = (xn > lo) && (xn < hi);
        if (fabs(dx) <= tol_abs + tol_rel * fabs(x)) {
            // converged: the evaluation at x is (to <=1e-14) the evaluation at the root
            if (inside) x = xn;
            converged = true;
            break;
        }
        if (hi - lo <= tol_abs + tol_rel * fabs(x)) { converged = true; break; }
        const bool shrinking = fabs(f) < T(0.75) * fprev;
        fprev = fabs(f);
        x = (inside && shrinking && finite_(xn)) ? xn : T(0.5) * (lo + hi);
    }
    if (type == JF_INV_ISIGMOID) {
        logd = log(v.Sp / (v.Sc * (v.Ss + v.ex))) + v.ex;
    } else {
        T y;
        inv_stage(type, v, y, logd);
    }
    logd_out = logd;
    // reference semantics: report elements whose residual exceeds 1e-7 (fp64) / 1e-4 (fp32)
    if (!(fabs(f) <= Num<T>::target_prec)) converged = false;
    return x;
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
    // normalisation needs the sum over k for each dimension
    for (int j = tid; j < d; j += nthreads) {
        T nmax = -Num<T>::big;
        if (c.norm_mode == JF_NORM_RAW)
            for (int k = 0; k < K; ++k) nmax = tmax(nmax, raw[c.raw_n() + k * d + j]);
        T sum = 0;
        for (int k = 0; k < K; ++k) {
            T g;
            if (c.norm_mode == JF_NORM_REGULATED) g = regulate_norm(raw[c.raw_n() + k * d + j], c.n_min, c.n_max);
            else if (c.norm_mode == JF_NORM_RAW) g = exp(raw[c.raw_n() + k * d + j] - nmax);
            else g = T(1);
            tab[c.tab_n() + k * d + j] = g;
            sum += g;
        }
        const T inv = T(1) / sum;
        for (int k = 0; k < K; ++k) tab[c.tab_n() + k * d + j] *= inv;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One layer on one row.  `p`: this row's raw parameter slice base (element i at p[i*sj]); `tab`: processed table.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int DM, int KM>
JF_DEVINL void gf_layer_logpdf(T* x, T& logdet, const GfLayerC<T>& c, int d, int K, bool processed, const T* tab,
                               const T* p, int64_t sj) {
    if (c.has_offset) {
#pragma unroll
        for (int j = 0; j < d; ++j) x[j] -= processed ? tab[c.tab_off + j] : p[(int64_t)(c.raw_off + j) * sj];
    }
    if (c.hh_iter > 0)
        householder_apply<T, DM>(x, d, c.hh_iter, true, processed, tab + c.tab_hh(), p + (int64_t)c.raw_hh() * sj, sj);
    T ld = 0;
#pragma unroll 1
    for (int j = 0; j < d; ++j) {
        Mix<T, KM> mx;
        load_mix<T, KM>(mx, c, K, j, processed, tab, p, sj);
        const MixVal<T> v = mix_eval<T, KM>(mx, K, x[j]);
        T y, logd;
        inv_stage(c.inv_type, v, y, logd);
        x[j] = y;
        ld += logd;
    }
    logdet += ld;
}

template <typename T, int DM, int KM>
JF_DEVINL void gf_layer_sample(T* x, T& logdet, const GfLayerC<T>& c, int d, int K, bool processed, const T* tab,
                               const T* p, int64_t sj, int& n_evals, int& n_unconv) {
    T ld = 0;
#pragma unroll 1
    for (int j = 0; j < d; ++j) {
        Mix<T, KM> mx;
        load_mix<T, KM>(mx, c, K, j, processed, tab, p, sj);
        T logd;
        int ev;
        bool conv;
        x[j] = gf_solve<T, KM>(mx, K, c.inv_type, x[j], logd, ev, conv);
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

}  // namespace jf
