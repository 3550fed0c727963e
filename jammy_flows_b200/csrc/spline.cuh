// Monotone rational-quadratic splines (Durkan et al. 2019) as the reference configures them for the "r" and "o" layers
// and for the vertical / circular sub-flows of "f": device math, one thread per row, everything in registers / local
// arrays.  What is computed (knot construction, closed-form smooth derivatives, bin search, forward map, quadratic-root
// inverse, log-derivative) follows the reference line by line because the clamps and the order of the floating-point
// operations are part of the numerical contract:
//   layers/spline_fns.py:13-19 (bin search), :45-186 (plain), :361-559 (smooth, C2), :561-760 (smooth circular)
//   layers/intervals/rational_quadratic_spline.py:180-246 and layers/spheres/splines_1d.py:111-170 (parameter unpacking)
// HOW is B200-first: no [B,n] knot tables, no gathers -- a thread builds the n+1 knots of its own row from the raw
// parameters (coalesced param-major loads), finds its bin while accumulating, and keeps only that bin's six numbers.
#pragma once
#include "common.cuh"

namespace jf {

template <typename T>
struct SplineC {
    int kind;        // JF_SPLINE_*
    int n_bins;
    int n_w, n_h, n_d;   // raw width / height / derivative parameters, stored in this order
    int fix_first, fix_second, indep;
    int bd_mode;     // JF_BD_*
    int natural_direction;
    int raw_off;     // offset of the spline's parameters from the start of the LAYER slice
    int pad_;
    T lo, hi, min_w, min_h, min_d, bd_fixed;
    T ln_max_ratio;  // (log(max_ratio) - log(n_bins-1))/2 when the width/height ratio is restricted, else <= 0
    __host__ __device__ int n_params() const { return n_w + n_h + n_d; }
};

// torch F.softplus with the default threshold 20 (part of the contract, like in the "g" layer)
template <typename T> JF_DEVINL T softplus_t(T t) { return t > T(20) ? t : log1p(exp(t)); }

// raw [n] -> knots k[0..n] on [lo,hi]: softmax with a floor, cumulated sequentially, ends pinned (spline_fns.py:86-98)
template <typename T>
JF_DEVINL void make_knots(T* raw, int n, T lo, T hi, T floor_, T ln_max_ratio, T* k) {
    if (ln_max_ratio > T(0)) {
        for (int i = 0; i < n; ++i) raw[i] = T(2) * (T(1) / (T(1) + exp(-raw[i]))) * ln_max_ratio - ln_max_ratio;
    }
    T m = raw[0];
    for (int i = 1; i < n; ++i) m = tmax(m, raw[i]);
    T sum = 0;
    for (int i = 0; i < n; ++i) { raw[i] = exp(raw[i] - m); sum += raw[i]; }
    const T rest = T(1) - floor_ * T(n);
    T cum = 0;
    k[0] = lo;
    for (int i = 0; i < n; ++i) {
        cum += floor_ + rest * (raw[i] / sum);
        k[i + 1] = (hi - lo) * cum + lo;
    }
    k[n] = hi;
}

// Evaluate the spline through knots (kx, ky) with knot derivatives d at x.  inverse: solve for the pre-image and return
// MINUS the forward log-derivative there (what the reference returns).  spline_fns.py:113-186.
template <typename T>
JF_DEVINL void rq_eval(const T* kx, const T* ky, const T* d, int n, bool inverse, T x, T& out, T& lad) {
    const T* ks = inverse ? ky : kx;
    int idx = -1;
    for (int i = 0; i < n; ++i) idx += (x >= ks[i]) ? 1 : 0;
    idx += (x >= ks[n] + T(1e-6)) ? 1 : 0;
    idx = idx < 0 ? 0 : (idx > n - 1 ? n - 1 : idx);   // outside the support: flagged by the caller, keep the loads valid
    const T x0 = kx[idx], wk = kx[idx + 1] - kx[idx];
    const T y0 = ky[idx], hk = ky[idx + 1] - ky[idx];
    const T s = hk / wk;
    const T d0 = d[idx], d1 = d[idx + 1];
    const T t = d0 + d1 - T(2) * s;
    T xi;
    if (inverse) {
        const T dy = x - y0;
        const T qa = dy * t + hk * (s - d0);
        const T qb = hk * d0 - dy * t;
        const T qc = -s * dy;
        const T disc = qb * qb - T(4) * qa * qc;
        xi = (T(2) * qc) / (-qb - sqrt(disc));
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
}

// One spline transformation of one coordinate.  `p`: base of the LAYER's raw parameter slice (element i at p[i*sj]);
// `scale` multiplies every raw parameter that is read (the per-row window of the "f" circular sub-flow, fvm_2d.py:416-427).
// Returns the number of out-of-support inputs (0/1); those are clamped into [lo,hi] (the reference raises instead).
template <typename T>
__device__ __noinline__ int spline_apply(const SplineC<T>& c, const T* p, int64_t sj, T scale, bool inverse, T x,
                                         T& out, T& lad) {
    T uw[JF_MAX_BINS + 1], uh[JF_MAX_BINS + 1], kx[JF_MAX_BINS + 1], ky[JF_MAX_BINS + 1], d[JF_MAX_BINS + 1];
    const int n = c.n_bins;
    const T* q = p + (int64_t)c.raw_off * sj;
    auto P = [&](int i) { return q[(int64_t)i * sj] * scale; };
    // ---- unpack (rational_quadratic_spline.py:207-246 / splines_1d.py:141-160) ----
    const bool mirror = (c.kind == JF_SPLINE_SMOOTH) && (n == 3);
    const int L = n - (mirror ? 1 : 0);
    const int zw = c.fix_first ? (c.fix_second ? 2 : 1) : 0, zh = c.fix_first ? 1 : 0;
    for (int i = 0; i < L; ++i) {
        uw[i] = i < zw ? T(0) : P(i - zw);
        uh[i] = i < zh ? T(0) : P(c.n_w + i - zh);
        if (c.indep) uh[i] += uw[i];
    }
    if (mirror) { uw[L] = uw[0]; uh[L] = uh[0]; }
    make_knots(uw, n, c.lo, c.hi, c.min_w, c.ln_max_ratio, kx);
    make_knots(uh, n, c.lo, c.hi, c.min_h, c.ln_max_ratio, ky);
    const int od = c.n_w + c.n_h;
    int oor = 0;
    if (c.kind == JF_SPLINE_PLAIN) {
        for (int i = 0; i <= n; ++i) {
            T raw;
            if (c.bd_mode == JF_BD_FIXED) raw = (i == 0 || i == n) ? c.bd_fixed : P(od + i - 1);
            else if (c.bd_mode == JF_BD_PERIODIC) raw = P(od + (i == n ? 0 : i));
            else raw = P(od + i);
            d[i] = c.min_d + softplus_t(raw);
        }
    } else if (c.kind == JF_SPLINE_SMOOTH) {
        const T b0 = c.min_d + softplus_t(c.bd_mode == JF_BD_FIXED ? c.bd_fixed : P(od + 0));
        const T b1 = c.min_d + softplus_t(c.bd_mode == JF_BD_FIXED ? c.bd_fixed : P(od + 1));
        if (n == 1) {
            d[0] = b0; d[1] = b1;
        } else if (n == 2) {                                   // spline_fns.py:433-453
            const T w1 = kx[1] - kx[0], w2 = kx[2] - kx[1], h1 = ky[1] - ky[0], h2 = ky[2] - ky[1];
            const T hs = h1 + h2;
            const T half = T(0.5) * ((h1 / hs) * (h2 / w2 - b1) + (h2 / hs) * (h1 / w1 - b0));
            const T qq = -(h1 * h2) * ((h1 / hs) * (T(1) / (w1 * w1)) + (h2 / hs) * (T(1) / (w2 * w2)));
            d[0] = b0; d[1] = half + sqrt(half * half - qq); d[2] = b1;
        } else {                                               // n == 3, symmetric: spline_fns.py:455-478
            const T w1 = kx[1] - kx[0], w2 = kx[2] - kx[1], h1 = ky[1] - ky[0], h2 = ky[2] - ky[1];
            const T cd = w1 * w2 * (T(2) * h1 + h2);
            const T pp = h2 * (b0 * w1 * w2 - h1 * (w1 + w2)) / cd;
            const T qq = -h1 * h2 * (h1 * w2 * w2 + h2 * w1 * w1) / (cd * w1 * w2);
            const T nh = -pp / T(2);
            const T mid = nh + sqrt(nh * nh - qq);
            d[0] = b0; d[1] = mid; d[2] = mid; d[3] = b1;
        }
    } else {                                                   // smooth circular, 2 bins: spline_fns.py:618-668
        const T two_pi = T(2 * kPi);
        const T w1 = kx[1] - kx[0], w2 = kx[2] - kx[1], h1 = ky[1] - ky[0], h2 = ky[2] - ky[1];
        const T hp = h1 * h2, wp = w1 * w2;
        const T a1 = h2 * w1, a2 = h1 * w2, ws = w1 + w2;
        const T root = sqrt(hp * (T(8) * (a1 * a1 + a2 * a2) + (T(9) * ws * ws - T(16) * wp) * hp));
        const T res = (hp * ws + root) / (T(4) * (h1 + h2) * wp);
        d[0] = res; d[1] = res; d[2] = res;
        const T a = -T(kPi) + w1 / T(2);
        const T ab = a + w2;
        const T nom = h2 * a * (a * h1 - res * w1 * ab);
        const T den = h1 * w2 * w2 + T(2) * (h1 - res * w1) * a * ab;
        const T corr = two_pi - (h1 + nom / den);
        const T mid_shift = T(kPi) - w1 / T(2);
        if (x < T(0) || x > two_pi) { oor = 1; x = clampv(x, T(0), two_pi); }
        T u = x - (inverse ? corr : mid_shift);
        if (u < T(0)) u += two_pi;
        rq_eval(kx, ky, d, n, inverse, u, out, lad);
        out += inverse ? mid_shift : corr;
        if (out > two_pi) out -= two_pi;
        if (x == T(0)) out = T(0);
        if (x == two_pi) out = two_pi;
        return oor;
    }
    if (x < c.lo || x > c.hi) { oor = 1; x = clampv(x, c.lo, c.hi); }
    rq_eval(kx, ky, d, n, inverse, x, out, lad);
    return oor;
}

}  // namespace jf
