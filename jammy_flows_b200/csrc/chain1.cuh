// One-dimensional sub-pdfs: intervals ("r") and the circle S1 ("o", "m"), all layers of the sub-pdf fused, one thread
// per row.  Reference (operation order and clamps kept literally):
//   interval chart      layers/intervals/interval_base.py:33-79
//   "r"                 layers/intervals/rational_quadratic_spline.py:180-400
//   S1 chart / rotation layers/spheres/sphere_base.py:460-480, :529-539, :601-695, :222-266
//   "o"                 layers/spheres/splines_1d.py:111-306
//   "m"                 layers/spheres/moebius_1d.py:57-259 (+ bisection_n_newton.py:137-256 for the non-natural direction)
#pragma once
#include "spline.cuh"
#include "subpdf_args.cuh"

namespace jf {

template <typename T>
struct Layer1C {
    int kind;          // JF_LAYER_RQS / JF_LAYER_S1SPLINE / JF_LAYER_MOEBIUS
    int first;         // layer carries the base chart of the sub-pdf
    int hh_iter;       // S1: Householder reflections in R^2 (0: no rotation); parameters come first in the slice
    int raw_off;       // start of the layer slice
    int K;             // m: number of Moebius components
    int natural_direction;
    T lo, hi;          // r: interval
    SplineC<T> sp;     // r, o
};

template <typename T>
struct Chain1Args {
    SubPdfArgs<T> a;
    int manifold;      // 'i' or 's'
    Layer1C<T> layers[JF_MAX_LAYERS];
};

constexpr int kMaxMoebius = 16;

// ---- S1 helpers ------------------------------------------------------------------------------------------------------
template <typename T>
JF_DEVINL T s1_rotate(T x, int n_iter, bool transpose, const T* p, int64_t sj) {
    T sx, cx;
    sincos(x, &sx, &cx);
    T e0 = cx, e1 = sx;
    for (int ii = 0; ii < n_iter; ++ii) {
        const int i = transpose ? ii : (n_iter - 1 - ii);
        const T v0 = p[(int64_t)(i * 2 + 0) * sj], v1 = p[(int64_t)(i * 2 + 1) * sj];
        const T c = T(2) * (v0 * e0 + v1 * e1) / (v0 * v0 + v1 * v1);
        e0 = fma(-c, v0, e0);
        e1 = fma(-c, v1, e1);
    }
    T arg = e0 / sqrt(e0 * e0 + e1 * e1);
    T a = acos(clampv(arg, T(-1), T(1)));
    return e1 < T(0) ? T(2 * kPi) - a : a;
}

// sphere -> line (sphere_base.py:460-480): fold to [0,pi] with a sign, z = sign*sqrt2*erfinv(1 - theta'/pi).
// erfinv(1-u) is evaluated as erfcinv(u): same function, no cancellation for small u ("Phi^-1").
template <typename T>
JF_DEVINL T s1_to_line(T x, T& logdet) {
    const bool neg = x > T(kPi);
    T y = neg ? T(2 * kPi) - x : x;
    const T eps = Prec<T>::f64 ? T(1e-8) : T(1e-5);
    if (y <= T(0)) y = eps;
    if (y >= T(2 * kPi)) y = T(2 * kPi) - eps;
    const T z = T(1.4142135623730951) * erfcinv(y / T(kPi));
    logdet += -T(kLogSqrt2Pi) + T(0.5) * z * z;
    return neg ? -z : z;
}

// line -> sphere (sphere_base.py:364-380, :529-539)
template <typename T>
JF_DEVINL T line_to_s1(T z, T& logdet) {
    const T r = fabs(z);
    logdet += T(kLogSqrt2Pi) - T(0.5) * r * r;
    const T a = T(kPi) * erfc(r * T(0.70710678118654752));     // pi*(1 - erf(r/sqrt2))
    return z >= T(0) ? a : T(2 * kPi) - a;
}

// ---- Moebius ----------------------------------------------------------------------------------------------------------
template <typename T>
struct MoebiusRow {
    T ox[kMaxMoebius], oy[kMaxMoebius], om[kMaxMoebius], l2[kMaxMoebius], cr[kMaxMoebius], sr[kMaxMoebius], wt[kMaxMoebius];
    int K;
};

// per-row constants of the K Moebius maps (moebius_1d.py:148-186): omega, the rotation that pins -pi -> -pi, weights
template <typename T>
JF_DEVINL void moebius_setup(MoebiusRow<T>& m, int K, const T* p, int64_t sj) {
    m.K = K;
    T nmax = -Num<T>::big;
    for (int k = 0; k < K; ++k) nmax = tmax(nmax, p[(int64_t)(k * 4 + 3) * sj]);
    T nsum = 0;
    for (int k = 0; k < K; ++k) {
        const T p0 = p[(int64_t)(k * 4 + 0) * sj], p1 = p[(int64_t)(k * 4 + 1) * sj], p2 = p[(int64_t)(k * 4 + 2) * sj];
        // 0.001 + exp(log(0.998) - logsumexp(0, -p2)) = 0.001 + 0.998*sigmoid(p2)
        const T sg = p2 >= T(0) ? T(1) / (T(1) + exp(-p2)) : exp(p2) / (T(1) + exp(p2));
        const T len = T(0.001) + T(0.998) * sg;
        const T inv = len / sqrt(p0 * p0 + p1 * p1);
        const T ox = p0 * inv, oy = p1 * inv;
        const T om = T(1) - len * len;
        // image of -pi (cos = -1, sin = sin(-pi) = -1.2246e-16 in the reference's numpy constants)
        const T cmp = T(-1), smp = T(-1.2246467991473532e-16);
        const T opo = T(1) + len * len - T(2) * (cmp * ox + smp * oy);
        const T rot = -T(kPi) - atan2(om * (smp - oy) - oy * opo, om * (cmp - ox) - ox * opo);
        T s, c;
        sincos(rot, &s, &c);
        m.ox[k] = ox; m.oy[k] = oy; m.om[k] = om; m.l2[k] = len * len; m.cr[k] = c; m.sr[k] = s;
        m.wt[k] = exp(p[(int64_t)(k * 4 + 3) * sj] - nmax);
        nsum += m.wt[k];
    }
    for (int k = 0; k < K; ++k) m.wt[k] /= nsum;
}

// value and derivative of the weighted Moebius map at x in [-pi,pi] (moebius_1d.py:140-259)
template <typename T>
JF_DEVINL void moebius_eval(const MoebiusRow<T>& m, T x, T& val, T& der) {
    T sx, cx;
    sincos(x, &sx, &cx);
    T v = 0, dsum = 0;
    for (int k = 0; k < m.K; ++k) {
        const T opo = T(1) + m.l2[k] - T(2) * (cx * m.ox[k] + sx * m.oy[k]);
        const T yv = m.om[k] * (sx - m.oy[k]) - m.oy[k] * opo;
        const T xv = m.om[k] * (cx - m.ox[k]) - m.ox[k] * opo;
        const T xp = m.cr[k] * xv - m.sr[k] * yv;
        const T yp = m.sr[k] * xv + m.cr[k] * yv;
        v = fma(m.wt[k], atan2(yp, xp) + T(kPi), v);
        dsum = fma(m.wt[k], m.om[k] / opo, dsum);
    }
    val = v - T(kPi);
    der = dsum;
}

// root of moebius(x) = target on [-pi,pi]: bracketed Newton (the map is strictly increasing); the reference runs 20
// bisections + <= 20 Newton steps to |dx| < 1e-14 (bisection_n_newton.py:137-256) and lands on the same root.
template <typename T>
JF_DEVINL T moebius_solve(const MoebiusRow<T>& m, T target, T& der_out, int& evals, bool& converged) {
    T lo = -T(kPi), hi = T(kPi), x = clampv(target, lo, hi);
    const T tol_abs = Num<T>::newton_abs_tol, tol_rel = T(4) * Num<T>::eps;
    T fprev = Num<T>::big, f = 0, der = 1;
    converged = false;
    evals = 0;
#pragma unroll 1
    for (int it = 0; it < 80; ++it) {
        T val;
        moebius_eval(m, x, val, der);
        ++evals;
        f = val - target;
        if (f < T(0)) lo = x; else hi = x;
        const T dx = f / der;
        const T xn = x - dx;
        const bool inside = (xn >= lo) && (xn <= hi);
        if (fabs(dx) <= tol_abs + tol_rel * fabs(x)) {
            if (inside) { x = xn; moebius_eval(m, x, val, der); ++evals; f = val - target; }
            converged = true;
            break;
        }
        if (hi - lo <= tol_abs + tol_rel * fabs(x)) { converged = true; break; }
        const bool shrinking = fabs(f) < T(0.75) * fprev;
        fprev = fabs(f);
        x = (inside && shrinking && finite_(xn)) ? xn : T(0.5) * (lo + hi);
    }
    if (!(fabs(f) <= Num<T>::target_prec)) converged = false;
    der_out = der;
    return x;
}

// the layer-intrinsic part of an S1 layer; logpdf = reference `_inv_flow_mapping`, else `_flow_mapping`
template <typename T>
__device__ __noinline__ T s1_inner(const Layer1C<T>& c, bool logpdf, T x, T& logdet, const T* pl, int64_t sj,
                                   int& oor, int& evals, int& unconv) {
    if (c.kind == JF_LAYER_S1SPLINE) {
        T out, lad;
        if (logpdf) {
            x = clampv(x, T(1e-7), T(2 * kPi - 1e-7));                      // splines_1d.py:120
            oor += spline_apply<T>(c.sp, pl, sj, T(1), c.natural_direction != 0, x, out, lad);
            out = clampv(out, T(1e-7), T(2 * kPi - 1e-7));
        } else {
            x = clampv(x, T(0), T(2 * kPi));                                // splines_1d.py:213-214
            oor += spline_apply<T>(c.sp, pl, sj, T(1), c.natural_direction == 0, x, out, lad);
            out = clampv(out, T(0), T(2 * kPi));
        }
        logdet += lad;
        return out;
    }
    // Moebius
    MoebiusRow<T> m;
    moebius_setup(m, c.K, pl, sj);
    if (x > T(kPi)) x -= T(2 * kPi);
    const bool direct = logpdf ? (c.natural_direction == 0) : (c.natural_direction != 0);
    T val, der;
    if (direct) {
        moebius_eval(m, x, val, der);
        logdet += log(der);
    } else {
        int ev; bool conv;
        val = moebius_solve(m, x, der, ev, conv);
        logdet -= log(der);
        evals += ev;
        unconv += conv ? 0 : 1;
    }
    return val < T(0) ? T(2 * kPi) + val : val;
}

template <typename T>
JF_DEVINL T layer1_logpdf(const Layer1C<T>& c, int manifold, T x, T& logdet, const T* p, int64_t sj, int& oor,
                          int& evals, int& unconv) {
    const T* pl = p + (int64_t)c.raw_off * sj;
    if (manifold == 'i') {
        T out, lad;
        x = clampv(x, T(-1), T(1));                                        // rational_quadratic_spline.py:297-298 (sic)
        oor += spline_apply<T>(c.sp, pl, sj, T(1), true, x, out, lad);
        logdet += lad;
        out = clampv(out, T(-1), T(1));
        if (c.first) {                                                       // interval_base.py:47-59
            const T width = c.hi - c.lo;
            const T u = (out - c.lo) / width;
            // sqrt2*erfinv(2u-1), evaluated from the smaller tail
            const T z = T(1.4142135623730951) * (u <= T(0.5) ? -erfcinv(T(2) * u) : erfcinv(T(2) * (T(1) - u)));
            logdet -= (-T(0.5) * z * z - T(kLogSqrt2Pi) + log(width));
            out = z;
        }
        return out;
    }
    if (c.hh_iter > 0) x = s1_rotate(x, c.hh_iter, true, pl, sj);
    x = s1_inner<T>(c, true, x, logdet, pl + (int64_t)(c.hh_iter * 2) * sj, sj, oor, evals, unconv);
    if (c.first) x = s1_to_line(x, logdet);
    return x;
}

template <typename T>
JF_DEVINL T layer1_sample(const Layer1C<T>& c, int manifold, T x, T& logdet, const T* p, int64_t sj, int& oor,
                          int& evals, int& unconv) {
    const T* pl = p + (int64_t)c.raw_off * sj;
    if (manifold == 'i') {
        if (c.first) {                                                       // interval_base.py:33-45
            const T width = c.hi - c.lo;
            logdet += -T(0.5) * x * x - T(kLogSqrt2Pi) + log(width);
            x = (T(0.5) + T(0.5) * erf(x * T(0.70710678118654752))) * width + c.lo;
        }
        T out, lad;
        x = clampv(x, T(-1), T(1));
        oor += spline_apply<T>(c.sp, pl, sj, T(1), false, x, out, lad);
        logdet += lad;
        return clampv(out, T(-1), T(1));
    }
    if (c.first) x = line_to_s1(x, logdet);
    x = s1_inner<T>(c, false, x, logdet, pl + (int64_t)(c.hh_iter * 2) * sj, sj, oor, evals, unconv);
    if (c.hh_iter > 0) x = s1_rotate(x, c.hh_iter, false, pl, sj);
    return x;
}

template <typename T, int DIR>
__global__ void __launch_bounds__(128) chain1_kernel(const __grid_constant__ Chain1Args<T> g) {
    const SubPdfArgs<T>& a = g.a;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.B) return;
    T x = a.in[row * a.ld_in];
    T logdet = a.logdet_in ? a.logdet_in[row] : T(0);
    const T* prow = a.params + row * a.sr;
    int oor = 0, evals = 0, unconv = 0;
    T z, target;
    if (DIR == JF_DIR_LOGPDF) {
        target = x;
        for (int l = a.n_layers - 1; l >= 0; --l)
            x = layer1_logpdf<T>(g.layers[l], g.manifold, x, logdet, prow, a.sj, oor, evals, unconv);
        z = x;
    } else {
        z = x;
        for (int l = 0; l < a.n_layers; ++l)
            x = layer1_sample<T>(g.layers[l], g.manifold, x, logdet, prow, a.sj, oor, evals, unconv);
        target = x;
    }
    if (a.emb_out) {
        if (g.manifold == 'i') {
            a.emb_out[row * a.ld_emb] = target;
        } else {
            T s, c;
            sincos(target, &s, &c);
            a.emb_out[row * a.ld_emb + 0] = c;
            a.emb_out[row * a.ld_emb + 1] = s;
        }
    }
    a.out[row * a.ld_out] = x;
    if (!finite_(x) || !finite_(logdet)) status_add(a.status, JF_STATUS_NONFINITE, 1);
    if (oor) status_add(a.status, JF_STATUS_OUT_OF_RANGE, oor);
    if (unconv) status_add(a.status, JF_STATUS_UNCONVERGED, unconv);
    status_add_warp(a.status, JF_STATUS_ITERATIONS, evals);
    if (a.logdet_out) a.logdet_out[row] = logdet;
    if (a.logbase_out) {
        const T prev = a.logbase_in ? a.logbase_in[row] : T(0);
        a.logbase_out[row] = prev - T(0.5) * z * z - T(kLogSqrt2Pi);
    }
}

}  // namespace jf
