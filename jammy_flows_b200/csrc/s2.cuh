// S2 charts and the Fisher-von-Mises layer "f" (reference defaults): device math, one thread per row.
// The operation ORDER (including the acos/cos round trips and every safety clamp) follows the reference literally,
// because the clamps are part of the numerical contract (SURVEY.md section 7 "quirks that must be replicated").
#pragma once
#include "common.cuh"

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

struct FvmLayerC {
    int add_rotation, hh_iter, first, raw_off;
    double z_sign, min_kappa;
};

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

// log_pdf direction: reference sphere_base.py:601-650 + fvm_2d.py:273-500 (+ sphere_to_plane :496-513, :416-430)
template <typename T>
JF_DEVINL void fvm_logpdf(T& c0, T& c1, T& logdet, const FvmLayerC& c, const T* p, int64_t sj) {
    T theta = c0, phi = c1;
    const T* pl = p + (int64_t)c.raw_off * sj;
    const int n_hh = c.add_rotation ? c.hh_iter * 3 : 0;
    if (c.add_rotation) {
        T e[3];
        s2_to_embedding(theta, phi, e, logdet);
        s2_rotate(e, c.hh_iter, true, pl, sj);
        s2_from_embedding(e, theta, phi, logdet);
    }
    const T kappa = exp(pl[(int64_t)n_hh * sj]) + T(c.min_kappa);
    const T s = T(c.z_sign);
    const T ct = cos(theta);
    logdet += log(sin(safe_angle(theta)));
    const T safe_part = kappa < T(100) ? log(exp(T(2) * kappa) - T(1)) : T(2) * kappa;
    logdet += log(T(2) * kappa) + kappa * (s * ct + T(1)) - safe_part;
    const T em2k = exp(T(-2) * kappa);
    T ret = s * ((T(1) + em2k - T(2) * exp(kappa * (s * ct - T(1)))) / (T(-1) + em2k));
    if (kappa < Num<T>::kappa_identity) ret = ct;
    ret = safe_costheta(ret, Num<T>::safe_costheta);
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
JF_DEVINL void fvm_sample(T& c0, T& c1, T& logdet, const FvmLayerC& c, const T* p, int64_t sj) {
    T theta, phi;
    const T* pl = p + (int64_t)c.raw_off * sj;
    const int n_hh = c.add_rotation ? c.hh_iter * 3 : 0;
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
    const T kappa = exp(pl[(int64_t)n_hh * sj]) + T(c.min_kappa);
    const T s = T(c.z_sign);
    const T ct = cos(theta);
    logdet += log(sin(safe_angle(theta)));
    logdet -= log(kappa * s * ct + kappa / tanh(kappa));
    T ret = s * (T(1) + (T(1) / kappa) * log(T(0.5) * (T(1) + s * ct) + (T(0.5) - T(0.5) * s * ct) * exp(T(-2) * kappa)));
    if (kappa < Num<T>::kappa_identity) ret = ct;
    ret = safe_costheta(ret, Num<T>::safe_costheta);
    theta = acos(ret);
    logdet -= log(sin(safe_angle(theta)));
    if (c.add_rotation) {
        T e[3];
        s2_to_embedding(theta, phi, e, logdet);
        s2_rotate(e, c.hh_iter, false, pl, sj);
        s2_from_embedding(e, theta, phi, logdet);
    }
    c0 = theta;
    c1 = phi;
}

}  // namespace jf
