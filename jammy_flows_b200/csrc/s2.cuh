// S2 charts and the Fisher-von-Mises layer "f" (reference defaults): device math, one thread per row.
// The operation ORDER (including the acos/cos round trips and every safety clamp) follows the reference literally,
// because the clamps are part of the numerical contract (SURVEY.md section 7 "quirks that must be replicated").
#pragma once
#include "spline.cuh"

namespace jf {

// reference layers/spheres/sphere_base.py:8-19
template <typename T> JF_DEVINL T safe_angle(T x) { return clampv(x, T(1e-7), T(kPi - 1e-7)); }
// reference layers/spheres/sphere_base.py:21-38
template <typename T> JF_DEVINL T safe_costheta(T x, T margin) { return clampv(x, T(-1) + margin, T(1) - margin); }

// (theta,phi) -> (x,y,z), logdet += log sin(theta_safe).  reference sphere_base.py:305-332
template <typename T>
JF_DEVINL void s2_to_embedding(T theta, T phi, T* e, T& logdet) {
    theta = safe_angle(theta);
    T st, ct, sp, cp;
    sincos(theta, &st, &ct);
    sincos(phi, &sp, &cp);
    e[0] = st * cp;
    e[1] = st * sp;
    e[2] = ct;
    logdet += log(st);
}

// (x,y,z) -> (theta,phi), logdet -= log sin(theta_safe).  reference sphere_base.py:266-282
template <typename T>
JF_DEVINL void s2_from_embedding(const T* e, T& theta, T& phi, T& logdet) {
    const T r = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    theta = safe_angle(acos(e[2] / r));
    logdet -= log(sin(theta));
    T arg = e[0] / sqrt(e[0] * e[0] + e[1] * e[1]);
    arg = arg > T(1) ? T(1) : arg;
    arg = arg < T(-1) ? T(-1) : arg;
    phi = acos(arg);
    if (e[1] < T(0)) phi = T(2 * kPi) - phi;
}

// one layer of an S2 sub-pdf: "f" (JF_LAYER_FVM) or "v" (JF_LAYER_EXPMAP)
struct FvmLayerC {
    int kind, add_rotation, hh_iter, first, raw_off;
    int rot_mode, n_rot;                            // f: JF_ROT_* of sphere_base.py:112-240 and its parameter count
    int kappa_mode, kappa_clamp;                    // f: JF_KAPPA_* (fvm_2d.py:108-138, :289-330)
    int extra_rot;                                  // f: add_extra_rotation_inbetween (fvm_2d.py:381-402, :664-688)
    double identity_region;                         // f: boundary_cos_theta_identity_region (fvm_2d.py:404-436, :573-612)
    int v_first, n_vertical, c_first, n_circular;   // f: nested spline sub-flows (indices into S2Args::splines)
    int K, natural_direction, max_iter;             // v: components / direction / iteration cap
    int pot, pad_pot;                               // v: JF_POT_*
    double z_sign, min_kappa;
};

// azimuthal window of the circular sub-flow, fvm_2d.py:267-271
template <typename T> JF_DEVINL T fvm_window(T c) {
    const T c3 = c * c * c;
    return c <= T(0) ? (T(6) * c3 * c * c + T(15) * c3 * c + T(10) * c3 + T(1))
                     : (T(-6) * c3 * c * c + T(15) * c3 * c - T(10) * c3 + T(1));
}

// pass-through chains of the nested layers (reference: vertical_rqspline_flow / circular_rqspline_flow are `pdf`s used
// as pass-throughs, main/default.py:998-1031 reversed for log_pdf / :1482-1506 in order for sampling)
template <typename T>
JF_DEVINL T fvm_vertical(const FvmLayerC& c, const SplineC<T>* sp, bool logpdf, T ret, T& logdet, const T* pl, int64_t sj,
                         int& oor) {
    for (int ii = 0; ii < c.n_vertical; ++ii) {
        const int i = c.v_first + (logpdf ? (c.n_vertical - 1 - ii) : ii);
        T out, lad;
        ret = clampv(ret, T(-1), T(1));
        oor += spline_apply<T>(sp[i], pl, sj, T(1), logpdf, ret, out, lad);
        ret = clampv(out, T(-1), T(1));
        logdet += lad;
    }
    return ret;
}
template <typename T>
JF_DEVINL T fvm_circular(const FvmLayerC& c, const SplineC<T>* sp, bool logpdf, T angle, T scale, T& logdet, const T* pl,
                         int64_t sj, int& oor) {
    for (int ii = 0; ii < c.n_circular; ++ii) {
        const int i = c.c_first + (logpdf ? (c.n_circular - 1 - ii) : ii);
        T out, lad;
        if (logpdf) {
            angle = clampv(angle, T(1e-7), T(2 * kPi - 1e-7));
            oor += spline_apply<T>(sp[i], pl, sj, scale, sp[i].natural_direction != 0, angle, out, lad);
            angle = clampv(out, T(1e-7), T(2 * kPi - 1e-7));
        } else {
            angle = clampv(angle, T(0), T(2 * kPi));
            oor += spline_apply<T>(sp[i], pl, sj, scale, sp[i].natural_direction == 0, angle, out, lad);
            angle = clampv(out, T(0), T(2 * kPi));
        }
        logdet += lad;
    }
    return angle;
}

// Householder reflections in R^3 on e.  Rotation parameters come first in the layer slice (sphere_base.py:630/673).
template <typename T>
JF_DEVINL void s2_rotate(T* e, int n_iter, bool transpose, const T* p, int64_t sj) {
    for (int ii = 0; ii < n_iter; ++ii) {
        const int i = transpose ? ii : (n_iter - 1 - ii);
        const T v0 = p[(int64_t)(i * 3 + 0) * sj], v1 = p[(int64_t)(i * 3 + 1) * sj], v2 = p[(int64_t)(i * 3 + 2) * sj];
        const T c = T(2) * (v0 * e[0] + v1 * e[1] + v2 * e[2]) / (v0 * v0 + v1 * v1 + v2 * v2);
        e[0] = fma(-c, v0, e[0]);
        e[1] = fma(-c, v1, e[1]);
        e[2] = fma(-c, v2, e[2]);
    }
}

// 3x3 rotation matrix of the non-Householder modes (reference sphere_base.py:127-216): "angles" = three Givens rotations
// over the index pairs (0,1), (0,2), (1,2) multiplied from the left, "xyz" = the rotation that takes the z axis to the
// normalised parameter vector, "quaternion" = the rotation of the (unnormalised) quaternion (a, i, j, k)
template <typename T>
JF_DEVINL void s2_rot_matrix(int mode, const T* p, int64_t sj, T (&M)[3][3]) {
    if (mode == JF_ROT_ANGLES) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = i == j ? T(1) : T(0);
        const int pa[3] = {0, 0, 1}, pb[3] = {1, 2, 2};
        for (int ind = 0; ind < 3; ++ind) {
            T sn, cs;
            sincos(p[(int64_t)ind * sj], &sn, &cs);
            const int a = pa[ind], b = pb[ind];
            for (int j = 0; j < 3; ++j) {                                // M <- G M: rows a and b mix
                const T ra = M[a][j], rb = M[b][j];
                M[a][j] = cs * ra + sn * rb;
                M[b][j] = -sn * ra + cs * rb;
            }
        }
    } else if (mode == JF_ROT_XYZ) {
        T nx = p[0], ny = p[sj], nz = p[2 * sj];
        const T inv = T(1) / sqrt(nx * nx + ny * ny + nz * nz);
        nx *= inv; ny *= inv; nz *= inv;
        const T q = T(1) + nz;
        M[0][0] = T(1) - nx * nx / q; M[0][1] = -nx * ny / q;       M[0][2] = nx;
        M[1][0] = -nx * ny / q;       M[1][1] = T(1) - ny * ny / q; M[1][2] = ny;
        M[2][0] = -nx;                M[2][1] = -ny;                M[2][2] = nz;
    } else {
        const T a = p[0], i = p[sj], j = p[2 * sj], k = p[3 * sj];
        const T n2 = a * a + i * i + j * j + k * k;
        M[0][0] = T(1) - T(2) * (j * j + k * k) / n2; M[0][1] = T(2) * (i * j - a * k) / n2;       M[0][2] = T(2) * (i * k + j * a) / n2;
        M[1][0] = T(2) * (i * j + a * k) / n2;       M[1][1] = T(1) - T(2) * (i * i + k * k) / n2; M[1][2] = T(2) * (j * k - i * a) / n2;
        M[2][0] = T(2) * (i * k - j * a) / n2;       M[2][1] = T(2) * (j * k + i * a) / n2;       M[2][2] = T(1) - T(2) * (i * i + j * j) / n2;
    }
}

// rotation of the "f" layer in embedding space: Householder reflections (default) or one of the matrix modes;
// transpose = inverse rotation (log_pdf direction, sphere_base.py:621)
template <typename T>
JF_DEVINL void fvm_rotate(T* e, const FvmLayerC& c, bool transpose, const T* p, int64_t sj) {
    if (c.rot_mode == JF_ROT_HOUSEHOLDER) { s2_rotate(e, c.hh_iter, transpose, p, sj); return; }
    T M[3][3];
    s2_rot_matrix(c.rot_mode, p, sj, M);
    const T x0 = e[0], x1 = e[1], x2 = e[2];
    if (transpose) {
        e[0] = M[0][0] * x0 + M[1][0] * x1 + M[2][0] * x2;
        e[1] = M[0][1] * x0 + M[1][1] * x1 + M[2][1] * x2;
        e[2] = M[0][2] * x0 + M[1][2] * x1 + M[2][2] * x2;
    } else {
        e[0] = M[0][0] * x0 + M[0][1] * x1 + M[0][2] * x2;
        e[1] = M[1][0] * x0 + M[1][1] * x1 + M[1][2] * x2;
        e[2] = M[2][0] * x0 + M[2][1] * x1 + M[2][2] * x2;
    }
}

// concentration of the "f" layer (reference fvm_2d.py:108-138 and :289-330): from its own raw parameter (three
// link functions, optionally clamped from below at -5) or from the norm of the rotation parameters
template <typename T>
JF_DEVINL T fvm_kappa(const FvmLayerC& c, const T* pl, int64_t sj) {
    if (c.kappa_mode >= JF_KAPPA_MU) {
        const int i0 = (c.kappa_mode == JF_KAPPA_MU || c.kappa_mode == JF_KAPPA_MU_SQUARED) ? 0 : 1;
        T s2 = 0;
        for (int i = i0; i < i0 + 3; ++i) { const T v = pl[(int64_t)i * sj]; s2 = fma(v, v, s2); }
        return (c.kappa_mode == JF_KAPPA_MU || c.kappa_mode == JF_KAPPA_QUATVEC) ? sqrt(s2) : s2;
    }
    T raw = pl[(int64_t)c.n_rot * sj];
    if (c.kappa_mode == JF_KAPPA_LOG_BOUNDED) {
        const T spl = raw > T(20) ? raw : log1p(exp(raw));               // F.softplus (threshold 20); clamping at -5 never binds
        return exp(spl + log(T(c.min_kappa)));
    }
    if (c.kappa_clamp) raw = tmax(raw, T(-5));
    if (c.kappa_mode == JF_KAPPA_SOFTPLUS) return (raw > T(20) ? raw : log1p(exp(raw))) + T(c.min_kappa);
    return exp(raw) + T(c.min_kappa);
}

// the fixed rotation between the z-scaling and the spline sub-flows (add_extra_rotation_inbetween): cylinder -> angle ->
// embedding -> M = [[0,0,1],[0,1,0],[-1,0,0]] (its transpose in the log_pdf direction) -> angle -> cylinder, with the
// reference's chain of log-det updates (fvm_2d.py:381-402 / :664-688)
template <typename T>
JF_DEVINL void fvm_extra_rotation(T& ret, T& phi, T& logdet, bool transpose) {
    T theta = acos(ret);
    logdet -= log(sin(safe_angle(theta)));
    T e[3];
    s2_to_embedding(theta, phi, e, logdet);
    const T x0 = e[0], x2 = e[2];
    if (transpose) { e[0] = -x2; e[2] = x0; } else { e[0] = x2; e[2] = -x0; }
    s2_from_embedding(e, theta, phi, logdet);
    ret = cos(theta);
    logdet += log(sin(safe_angle(theta)));
}

// log_pdf direction: reference sphere_base.py:601-650 + fvm_2d.py:273-500 (+ sphere_to_plane :496-513, :416-430)
template <typename T>
JF_DEVINL void fvm_logpdf(T& c0, T& c1, T& logdet, const FvmLayerC& c, const SplineC<T>* sp, const T* p, int64_t sj,
                          int& oor) {
    T theta = c0, phi = c1;
    const T* pl = p + (int64_t)c.raw_off * sj;
    const int n_hh = c.n_rot + (c.kappa_mode >= JF_KAPPA_MU ? -1 : 0);   // (+1 below: the layer's own kappa parameter, if any)
    if (c.add_rotation) {
        T e[3];
        s2_to_embedding(theta, phi, e, logdet);
        fvm_rotate(e, c, true, pl, sj);
        s2_from_embedding(e, theta, phi, logdet);
    }
    const T kappa = fvm_kappa(c, pl, sj);
    const T s = T(c.z_sign);
    const T ct = cos(theta);
    logdet += log(sin(safe_angle(theta)));
    const T safe_part = kappa < T(100) ? log(exp(T(2) * kappa) - T(1)) : T(2) * kappa;
    logdet += log(T(2) * kappa) + kappa * (s * ct + T(1)) - safe_part;
    const T em2k = exp(T(-2) * kappa);
    T ret = s * ((T(1) + em2k - T(2) * exp(kappa * (s * ct - T(1)))) / (T(-1) + em2k));
    if (kappa < Num<T>::kappa_identity) ret = ct;
    ret = safe_costheta(ret, T(Num<T>::safe_costheta));
    const T* psub = pl + (int64_t)(n_hh + 1) * sj;     // spline parameters follow kappa (fvm_2d.py:311-316)
    if (c.extra_rot) fvm_extra_rotation(ret, phi, logdet, true);
    // sub-flows act only inside the identity region's complement (all rows when the region is 0)
    const bool contained = c.identity_region == 0.0 || (ret > T(-1.0 + c.identity_region) && ret < T(1.0 - c.identity_region));
    if (contained) {
        if (c.n_circular > 0) phi = fvm_circular<T>(c, sp, true, phi, fvm_window(ret), logdet, psub, sj, oor);
        if (c.n_vertical > 0) ret = fvm_vertical<T>(c, sp, true, ret, logdet, psub, sj, oor);
    }
    ret = safe_costheta(ret, T(Num<T>::safe_costheta));
    theta = acos(ret);
    logdet -= log(sin(safe_angle(theta)));
    if (c.first) {
        // sphere -> plane (stereographic-Gaussian chart)
        const T th = safe_angle(theta);
        const T cx = safe_costheta(cos(th), T(1e-6));
        const T r = sqrt(-log((T(1) - cx) * T(0.5)) * T(2));
        logdet += -log(T(1) - cx) + log(sin(th));
        T sp, cp;
        sincos(phi, &sp, &cp);
        c0 = r * cp;
        c1 = r * sp;
    } else {
        c0 = theta;
        c1 = phi;
    }
}

// sampling direction: reference sphere_base.py:653-695 (+ plane_to_sphere :364-408, :569-592) + fvm_2d.py:502-726
template <typename T>
JF_DEVINL void fvm_sample(T& c0, T& c1, T& logdet, const FvmLayerC& c, const SplineC<T>* sp, const T* p, int64_t sj,
                          int& oor) {
    T theta, phi;
    const T* pl = p + (int64_t)c.raw_off * sj;
    const int n_hh = c.n_rot + (c.kappa_mode >= JF_KAPPA_MU ? -1 : 0);
    if (c.first) {
        const T r = sqrt(c0 * c0 + c1 * c1);
        const T arg = (r == T(0)) ? T(1) : c0 / r;
        phi = acos(arg);
        if (c1 < T(0)) phi = T(2 * kPi) - phi;
        theta = safe_angle(acos(T(1) - T(2) * exp(-(r * r) * T(0.5))));
        logdet += log(T(1) - cos(theta)) - log(sin(theta));
    } else {
        theta = c0;
        phi = c1;
    }
    const T kappa = fvm_kappa(c, pl, sj);
    const T s = T(c.z_sign);
    T ct = cos(theta);
    logdet += log(sin(safe_angle(theta)));
    const T* psub = pl + (int64_t)(n_hh + 1) * sj;
    const bool contained = c.identity_region == 0.0 || (ct > T(-1.0 + c.identity_region) && ct < T(1.0 - c.identity_region));
    if (contained) {
        if (c.n_vertical > 0) ct = fvm_vertical<T>(c, sp, false, ct, logdet, psub, sj, oor);                       // fvm_2d.py:591
        if (c.n_circular > 0) phi = fvm_circular<T>(c, sp, false, phi, fvm_window(ct), logdet, psub, sj, oor);     // :595-607
    }
    if (c.extra_rot) fvm_extra_rotation(ct, phi, logdet, false);
    logdet -= log(kappa * s * ct + kappa / tanh(kappa));
    T ret = s * (T(1) + (T(1) / kappa) * log(T(0.5) * (T(1) + s * ct) + (T(0.5) - T(0.5) * s * ct) * exp(T(-2) * kappa)));
    if (kappa < Num<T>::kappa_identity) ret = ct;
    ret = safe_costheta(ret, T(Num<T>::safe_costheta));
    theta = acos(ret);
    logdet -= log(sin(safe_angle(theta)));
    if (c.add_rotation) {
        T e[3];
        s2_to_embedding(theta, phi, e, logdet);
        fvm_rotate(e, c, false, pl, sj);
        s2_from_embedding(e, theta, phi, logdet);
    }
    c0 = theta;
    c1 = phi;
}

// ---------------------------------------------------------------------------------------------------------------------
// "v": exponential-map flow on S2 with the exponential potential
//   reference layers/spheres/exponential_map_s2.py:248-442 (map + Jacobian), :446-528 (directions),
//   layers/bisection_n_newton.py:394-465 (inverse by damped descent on the sphere, <= 1000 iterations)
// grad phi(x) = sum_k w_k mu_k exp(beta_k (x.mu_k - 1));  y = exp_x(grad phi projected on the tangent plane);
// log-det = 1/2 log det( (J B)^T (J B) ), J the 3x3 Jacobian of y in embedding space, B = [t, x cross t].
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxExpComp = 16;
constexpr int kVSplineBins = 10;        // exp_map_type "splines": num_spline_basis_functions (exponential_map_s2.py:111)

template <typename T>
struct VRow {
    T mx[kMaxExpComp], my[kMaxExpComp], mz[kMaxExpComp], w[kMaxExpComp], beta[kMaxExpComp];
    int K, pot;
    const T* sp_p;      // "splines" potential: raw spline parameters of component 0 (rows 4 .. 4 + 3 n + 1 of the [n_pot, K] block)
    int64_t sp_sj;      // stride between consecutive spline parameters of one component (K * sj)
    int64_t sp_ck;      // stride between components (sj)
};

// parameters [5, K] (index i*K + k): rows 0-2 mean direction (its length sets the weight bound), 3 log-weight, 4 log beta
template <typename T>
JF_DEVINL void v_setup(VRow<T>& r, int K, int pot, const T* p, int64_t sj) {
    r.K = K;
    r.pot = pot;
    r.sp_p = p + (int64_t)(4 * K) * sj;
    r.sp_sj = (int64_t)K * sj;
    r.sp_ck = sj;
    T lmax = -Num<T>::big;
    for (int k = 0; k < K; ++k) lmax = tmax(lmax, p[(int64_t)(3 * K + k) * sj]);
    T lsum = 0;
    for (int k = 0; k < K; ++k) lsum += exp(p[(int64_t)(3 * K + k) * sj] - lmax);
    const T lse = lmax + log(lsum);
    for (int k = 0; k < K; ++k) {
        const T a = p[(int64_t)(0 * K + k) * sj], b = p[(int64_t)(1 * K + k) * sj], c = p[(int64_t)(2 * K + k) * sj];
        const T n = sqrt(a * a + b * b + c * c);
        r.mx[k] = a / n; r.my[k] = b / n; r.mz[k] = c / n;
        const T fake = -log(T(1) + T(1.718281828459045) * exp(-n / T(10))) + T(1);          // :32-43, :262
        r.w[k] = exp(p[(int64_t)(3 * K + k) * sj] - lse + log(fake));                       // :288-289
        r.beta[k] = pot == JF_POT_EXPONENTIAL ? exp(p[(int64_t)(4 * K + k) * sj]) : T(0);   // :296
    }
}

// y = map(x), J = dy/dx (row-major 3x3), hld = 1/2 log|det((J B)^T (J B))|
template <typename T>
__device__ __noinline__ void v_eval(const VRow<T>& r, const T* x, T* y, T* J, T& hld) {
    T g0 = 0, g1 = 0, g2 = 0, G00 = 0, G01 = 0, G02 = 0, G11 = 0, G12 = 0, G22 = 0;
    SplineC<T> sc;
    if (r.pot == JF_POT_SPLINES) {
        // plain 10-bin spline on [-1,1] -> [-1,1] with (n, n, n+1) raw parameters, exponential_map_s2.py:358-366
        sc.kind = JF_SPLINE_PLAIN; sc.n_bins = kVSplineBins; sc.n_w = kVSplineBins; sc.n_h = kVSplineBins; sc.n_d = kVSplineBins + 1;
        sc.fix_first = 0; sc.fix_second = 0; sc.indep = 0; sc.bd_mode = JF_BD_PARAMS; sc.natural_direction = 0; sc.raw_off = 0;
        sc.pad_ = 0;
        sc.lo = T(-1); sc.hi = T(1); sc.min_w = T(1e-3); sc.min_h = T(1e-3); sc.min_d = T(1e-3); sc.bd_fixed = T(0);
        sc.ln_max_ratio = T(-1);
    }
    for (int k = 0; k < r.K; ++k) {
        const T xm = x[0] * r.mx[k] + x[1] * r.my[k] + x[2] * r.mz[k];
        // gradient of the potential and its Jacobian coefficient per component (exponential_map_s2.py:285-344)
        T c, cb;
        if (r.pot == JF_POT_EXPONENTIAL) { c = r.w[k] * exp(r.beta[k] * (xm - T(1))); cb = c * r.beta[k]; }
        else if (r.pot == JF_POT_QUADRATIC) { c = r.w[k] * xm; cb = r.w[k]; }
        else if (r.pot == JF_POT_SPLINES) {
            // the potential is the integral of the spline: its gradient is w mu s(x.mu), the Jacobian w mu mu^T s'(x.mu)
            T sv, lad;
            spline_apply<T>(sc, r.sp_p + (int64_t)k * r.sp_ck, r.sp_sj, T(1), false, xm, sv, lad);
            c = r.w[k] * sv;
            cb = r.w[k] * exp(lad);
        }
        else { c = r.w[k]; cb = T(0); }
        g0 = fma(c, r.mx[k], g0); g1 = fma(c, r.my[k], g1); g2 = fma(c, r.mz[k], g2);
        G00 = fma(cb, r.mx[k] * r.mx[k], G00); G01 = fma(cb, r.mx[k] * r.my[k], G01); G02 = fma(cb, r.mx[k] * r.mz[k], G02);
        G11 = fma(cb, r.my[k] * r.my[k], G11); G12 = fma(cb, r.my[k] * r.mz[k], G12); G22 = fma(cb, r.mz[k] * r.mz[k], G22);
    }
    const T g[3] = {g0, g1, g2};
    const T G[3][3] = {{G00, G01, G02}, {G01, G11, G12}, {G02, G12, G22}};
    // unnormalised logarithmic map of g at x and its Jacobians (:153-214)
    const T tn = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
    const T nh[3] = {g0 / tn, g1 / tn, g2 / tn};
    // (the dot product of two unit vectors can exceed 1 by a rounding error when the gradient is parallel to x: acos and
    //  sqrt(1 - ca^2) then return NaN -- once per ~4 M rows in fp32, never observed in fp64; the reference asserts fp64
    //  for this layer, exponential_map_s2.py:448.  Kept a few ulp inside (-1, 1): the map is the identity there.)
    const T ca_lim = T(1) - T(4) * Num<T>::eps;
    const T ca = clampv(nh[0] * x[0] + nh[1] * x[1] + nh[2] * x[2], -ca_lim, ca_lim);
    const T sa = sin(acos(ca));
    const T sq = sqrt(T(1) - ca * ca);
    T th[3], u[3];
    for (int i = 0; i < 3; ++i) {
        th[i] = (nh[i] - x[i] * ca) / sa;
        u[i] = (x[i] - nh[i] * ca) / (sa * sa);
    }
    const T proj = g[0] * th[0] + g[1] * th[1] + g[2] * th[2];
    T nG[3], DG[3][3], xDG[3], thG[3];
    for (int l = 0; l < 3; ++l) nG[l] = nh[0] * G[0][l] + nh[1] * G[1][l] + nh[2] * G[2][l];
    for (int i = 0; i < 3; ++i)
        for (int l = 0; l < 3; ++l) DG[i][l] = (G[i][l] - nh[i] * nG[l]) / tn;
    for (int l = 0; l < 3; ++l) {
        xDG[l] = x[0] * DG[0][l] + x[1] * DG[1][l] + x[2] * DG[2][l];
        thG[l] = th[0] * G[0][l] + th[1] * G[1][l] + th[2] * G[2][l];
    }
    const T cot = -ca / sa;
    const T gu = g[0] * u[0] + g[1] * u[1] + g[2] * u[2];
    T Jp[3], Jt[3][3];
    for (int l = 0; l < 3; ++l) {
        Jp[l] = cot * g[l] + gu * (-nh[l] / sq) + thG[l];
        const T rowv = (-nh[l] - xDG[l]) / sq;
        for (int i = 0; i < 3; ++i) Jt[i][l] = (i == l ? cot : T(0)) + u[i] * rowv + DG[i][l] / sa;
    }
    T sp, cp;
    sincos(proj, &sp, &cp);
    for (int i = 0; i < 3; ++i) {
        y[i] = x[i] * cp + th[i] * sp;
        for (int l = 0; l < 3; ++l)
            J[i * 3 + l] = (i == l ? cp : T(0)) + (-x[i] * sp) * Jp[l] + Jt[i][l] * sp + th[i] * cp * Jp[l];
    }
    const T t2[3] = {x[1] * th[2] - x[2] * th[1], x[2] * th[0] - x[0] * th[2], x[0] * th[1] - x[1] * th[0]};
    T a[3], b[3];
    for (int i = 0; i < 3; ++i) {
        a[i] = J[i * 3 + 0] * th[0] + J[i * 3 + 1] * th[1] + J[i * 3 + 2] * th[2];
        b[i] = J[i * 3 + 0] * t2[0] + J[i * 3 + 1] * t2[1] + J[i * 3 + 2] * t2[2];
    }
    const T aa = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], bb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    const T ab = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    hld = T(0.5) * log(fabs(aa * bb - ab * ab));
}

// orthonormal basis of the tangent plane at the unit vector x
template <typename T>
JF_DEVINL void tangent_basis(const T* x, T* e1, T* e2) {
    // cross with the coordinate axis least aligned with x
    const T ax = fabs(x[0]), ay = fabs(x[1]), az = fabs(x[2]);
    T h[3] = {T(0), T(0), T(0)};
    if (ax <= ay && ax <= az) h[0] = T(1); else if (ay <= az) h[1] = T(1); else h[2] = T(1);
    e1[0] = x[1] * h[2] - x[2] * h[1]; e1[1] = x[2] * h[0] - x[0] * h[2]; e1[2] = x[0] * h[1] - x[1] * h[0];
    const T n = T(1) / sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    e1[0] *= n; e1[1] *= n; e1[2] *= n;
    e2[0] = x[1] * e1[2] - x[2] * e1[1]; e2[1] = x[2] * e1[0] - x[0] * e1[2]; e2[2] = x[0] * e1[1] - x[1] * e1[0];
}

// Pre-image of `tg` under the map: Newton on the sphere in tangent coordinates with backtracking on the merit
// 1 - y(x).tg.  The map is a global diffeomorphism (Sei 2009), so this is the same point the reference's damped descent
// (step 0.4, stop at 1e-12, <= 1000 iterations) converges to; quadratic instead of linear convergence.
template <typename T>
__device__ __noinline__ void v_solve(const VRow<T>& r, const T* tg, int max_iter, T* x, T& hld, int& evals, bool& converged) {
    x[0] = tg[0]; x[1] = tg[1]; x[2] = tg[2];
    T y[3], J[9];
    v_eval(r, x, y, J, hld);
    evals = 1;
    converged = false;
    T merit = T(1) - (y[0] * tg[0] + y[1] * tg[1] + y[2] * tg[2]);
    const T tol = Prec<T>::f64 ? T(1e-13) : T(1e-6);
    const int cap = max_iter < 200 ? max_iter : 200;
    T prev_nrm = Num<T>::big;
#pragma unroll 1
    for (int it = 0; it < cap; ++it) {
        // residual: logarithmic map of tg at y, expressed in a tangent basis at y
        T f1[3], f2[3], e1[3], e2[3];
        tangent_basis(y, f1, f2);
        tangent_basis(x, e1, e2);
        const T cg = clampv(y[0] * tg[0] + y[1] * tg[1] + y[2] * tg[2], T(-1), T(1));
        T rv[3];
        {
            T d0 = tg[0] - y[0] * cg, d1 = tg[1] - y[1] * cg, d2 = tg[2] - y[2] * cg;
            const T dn = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            const T gamma = atan2(dn, cg);
            const T sc = dn > T(0) ? gamma / dn : T(0);
            rv[0] = d0 * sc; rv[1] = d1 * sc; rv[2] = d2 * sc;
        }
        const T r1 = f1[0] * rv[0] + f1[1] * rv[1] + f1[2] * rv[2], r2 = f2[0] * rv[0] + f2[1] * rv[1] + f2[2] * rv[2];
        T Je1[3], Je2[3];
        for (int i = 0; i < 3; ++i) {
            Je1[i] = J[i * 3 + 0] * e1[0] + J[i * 3 + 1] * e1[1] + J[i * 3 + 2] * e1[2];
            Je2[i] = J[i * 3 + 0] * e2[0] + J[i * 3 + 1] * e2[1] + J[i * 3 + 2] * e2[2];
        }
        const T A11 = f1[0] * Je1[0] + f1[1] * Je1[1] + f1[2] * Je1[2], A12 = f1[0] * Je2[0] + f1[1] * Je2[1] + f1[2] * Je2[2];
        const T A21 = f2[0] * Je1[0] + f2[1] * Je1[1] + f2[2] * Je1[2], A22 = f2[0] * Je2[0] + f2[1] * Je2[1] + f2[2] * Je2[2];
        const T det = A11 * A22 - A12 * A21;
        T dl1 = (A22 * r1 - A12 * r2) / det, dl2 = (A11 * r2 - A21 * r1) / det;
        T nrm = sqrt(dl1 * dl1 + dl2 * dl2);
        if (!finite_(nrm)) break;
        if (nrm > T(1)) { dl1 /= nrm; dl2 /= nrm; nrm = T(1); }
        // converged: the Newton step is below the tolerance, or it has stopped shrinking at a small length, i.e. it is
        // rounding noise of the residual (quadratic convergence shrinks it by orders of magnitude per step otherwise).
        // In fp32 the noise floor (~3e-6 for a Jacobian of condition ~10) sits ABOVE the fixed tolerance: without the
        // second test most rows ran into the iteration cap (measured: 780 evaluations per row on cfg4).
        if (nrm <= tol || (nrm < (Prec<T>::f64 ? T(1e-9) : T(1e-3)) && nrm >= T(0.5) * prev_nrm)) {
            converged = true;
            if (nrm == T(0)) break;
        }
        prev_nrm = nrm;
        // step along the geodesic, halving while the merit does not decrease
        T xn[3], yn[3], Jn[9], hn, mn = merit;
        bool ok = false;
        T scale = T(1);
        for (int bt = 0; bt < 12; ++bt) {
            const T len = nrm * scale;
            T sl, cl;
            sincos(len, &sl, &cl);
            for (int i = 0; i < 3; ++i) xn[i] = x[i] * cl + ((e1[i] * dl1 + e2[i] * dl2) / nrm) * sl;
            const T xnrm = T(1) / sqrt(xn[0] * xn[0] + xn[1] * xn[1] + xn[2] * xn[2]);
            xn[0] *= xnrm; xn[1] *= xnrm; xn[2] *= xnrm;
            v_eval(r, xn, yn, Jn, hn);
            ++evals;
            mn = T(1) - (yn[0] * tg[0] + yn[1] * tg[1] + yn[2] * tg[2]);
            // near the root 1 - y.tg is pure rounding noise (~1e-16): trust the Newton step there
            if (mn <= merit || converged || nrm < T(1e-5)) { ok = true; break; }
            scale *= T(0.5);
        }
        if (!ok) break;
        for (int i = 0; i < 3; ++i) { x[i] = xn[i]; y[i] = yn[i]; }
        for (int i = 0; i < 9; ++i) J[i] = Jn[i];
        hld = hn;
        merit = mn;
        if (converged) break;
    }
    if (!(merit <= (Prec<T>::f64 ? T(1e-10) : T(1e-5)))) converged = false;
}

// reference exponential_map_s2.py:446-528 wrapped by sphere_base.py:601-695
template <typename T>
JF_DEVINL void v_layer(bool logpdf, T& c0, T& c1, T& logdet, const FvmLayerC& c, const T* p, int64_t sj, int& evals,
                       int& unconv) {
    T theta, phi;
    const T* pl = p + (int64_t)c.raw_off * sj;
    const int n_hh = c.add_rotation ? c.hh_iter * 3 : 0;
    if (!logpdf && c.first) {
        const T r = sqrt(c0 * c0 + c1 * c1);
        const T arg = (r == T(0)) ? T(1) : c0 / r;
        phi = acos(arg);
        if (c1 < T(0)) phi = T(2 * kPi) - phi;
        theta = safe_angle(acos(T(1) - T(2) * exp(-(r * r) * T(0.5))));
        logdet += log(T(1) - cos(theta)) - log(sin(theta));
    } else {
        theta = c0; phi = c1;
    }
    T e[3];
    if (logpdf && c.add_rotation) {
        s2_to_embedding(theta, phi, e, logdet);
        s2_rotate(e, c.hh_iter, true, pl, sj);
        s2_from_embedding(e, theta, phi, logdet);
    }
    VRow<T> row;
    v_setup(row, c.K, c.pot, pl + (int64_t)n_hh * sj, sj);
    s2_to_embedding(theta, phi, e, logdet);
    T y[3], J[9], hld;
    const bool direct = logpdf ? (c.natural_direction == 0) : (c.natural_direction != 0);
    if (direct) {
        v_eval(row, e, y, J, hld);
        logdet += hld;
    } else {
        int ev; bool conv;
        v_solve(row, e, c.max_iter, y, hld, ev, conv);
        logdet -= hld;
        evals += ev;
        unconv += conv ? 0 : 1;
    }
    s2_from_embedding(y, theta, phi, logdet);
    if (!logpdf && c.add_rotation) {
        s2_to_embedding(theta, phi, e, logdet);
        s2_rotate(e, c.hh_iter, false, pl, sj);
        s2_from_embedding(e, theta, phi, logdet);
    }
    if (logpdf && c.first) {
        const T th = safe_angle(theta);
        const T cx = safe_costheta(cos(th), T(1e-6));
        const T r = sqrt(-log((T(1) - cx) * T(0.5)) * T(2));
        logdet += -log(T(1) - cx) + log(sin(th));
        T sp, cp;
        sincos(phi, &sp, &cp);
        c0 = r * cp; c1 = r * sp;
    } else {
        c0 = theta; c1 = phi;
    }
}

}  // namespace jf
