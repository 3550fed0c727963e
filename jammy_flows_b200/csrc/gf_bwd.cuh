// Element-level pieces of the backward of the Euclidean "g"-chain log_pdf (training, BASELINE configs[4]; the kernel
// that drives them is csrc/gf_fb.cuh): per-row gradient of
//   log p(x) = sum_j log N(z_j) + sum_layers sum_j log y'(v_j)
// with respect to the PER-ROW raw parameters the MLP emitted (the reference gets these from autograd through
// gaussianization_flow.py:995-1057, :699-861, :389-454; here they are closed-form, one thread per row, nothing stored
// in HBM but the inputs and the gradient buffer).
//
//   forward (recomputed): for l = L-1 .. 0:  u = x - offset_l ; v = Q_l^T u ; (y_j, l_j) = stage_l(v_j) ; x = y
//                         the pre-stage vectors v of every layer are kept in local memory (L*d values)
//   backward: ybar = -z * g_row (from the base density), lbar = g_row; for l = 0 .. L-1:
//       per dimension: closed-form derivatives of the K-logistic mixture and of the inverse-CDF stage with respect to
//           v, the means, 1/w and the weights, chained through the width / norm regulators to the raw parameters;
//       rotation: the Householder reflections are involutions, so their inputs are recovered by re-applying them to
//           the output -- no intermediate is stored; gradients with respect to the reflection vectors in closed form;
//       offset: minus the input gradient.
// Supported stages: isigmoid and inormal_partly_precise (bulk and Pade tails); per-row parameters only.
#pragma once
#ifndef JF_BWD_K_UNROLL
#define JF_BWD_K_UNROLL 1
#endif
#define JF_BWD_PRAGMA_(x) _Pragma(#x)
#define JF_BWD_PRAGMA(x) JF_BWD_PRAGMA_(x)
#define JF_BWD_UNROLL JF_BWD_PRAGMA(unroll JF_BWD_K_UNROLL)
#include "gf.cuh"
#include "subpdf_args.cuh"

namespace jf {

// regulated parameters of (layer, dimension j) into this thread's slots WITHOUT normalising the weights; returns their sum
template <typename T>
__device__ __noinline__ T bwd_regulate(const GfLayerC<T>& c, int K, int j, const T* p, int64_t sj, T* slots) {
    const int d = c.d, nt = blockDim.x;
    T* sm = slots + threadIdx.x;
    T* si = sm + (size_t)K * nt;
    T* sn = si + (size_t)K * nt;
    const T* pm = p + (int64_t)(c.raw_m() + j) * sj;
    const T* pw = p + (int64_t)(c.raw_w() + j) * sj;
    const T* pn = p + (int64_t)(c.raw_n() + j) * sj;
    const int64_t step = (int64_t)d * sj;
    // phase 1: raw values -> slots (a pure copy loop keeps 3 x 8 independent loads in flight: the latency of the parameter
    // stream is paid once per (layer, dimension), not once per kernel)
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        sm[(size_t)k * nt] = pm[k * step];
        si[(size_t)k * nt] = pw[k * step];
        sn[(size_t)k * nt] = (c.norm_mode != JF_NORM_NONE) ? pn[k * step] : T(0);
    }
    // phase 2: regulate in place
    T nmax = -Num<T>::big;
    if (c.norm_mode == JF_NORM_RAW)
        JF_BWD_UNROLL
        for (int k = 0; k < K; ++k) nmax = tmax(nmax, sn[(size_t)k * nt]);
    T G = 0;
    JF_BWD_UNROLL
    for (int k = 0; k < K; ++k) {
        si[(size_t)k * nt] = regulate_inv_width(si[(size_t)k * nt], c.w_min, c.inv_w_max);
        T g;
        if (c.norm_mode == JF_NORM_REGULATED) g = regulate_norm(sn[(size_t)k * nt], c.n_min, c.n_max);
        else if (c.norm_mode == JF_NORM_RAW) g = exp(sn[(size_t)k * nt] - nmax);
        else g = T(1);
        sn[(size_t)k * nt] = g;
        G += g;
    }
    return G;
}

// backward of one element: v -> (y, l).  gy, gl: upstream gradients of y and l.  Writes the raw-parameter gradients of
// this (layer, dimension) and returns the gradient with respect to v.  Returns false for rows outside the bulk.
template <typename T>
__device__ __noinline__ bool gf_elem_backward(const GfLayerC<T>& c, int K, int j, T v, T gy, T gl, T G, const T* slots,
                                              T* gp /* grad_params row base */, int64_t sj, T& vbar) {
    const int nt = blockDim.x;
    const T* sm = slots + threadIdx.x;
    const T* si = sm + (size_t)K * nt;
    const T* sn = si + (size_t)K * nt;
    const T invG = rcp_pos_(G);
    // ---- pass 1: common rescaling exponent (all kernels on one side of v) ----
    bool any_pos = false, any_neg = false;
    T delta = Num<T>::big;
    JF_BWD_UNROLL
    for (int k = 0; k < K; ++k) {
        const T a = (v - sm[(size_t)k * nt]) * si[(size_t)k * nt];
        any_pos = any_pos || (a >= T(0));
        any_neg = any_neg || (a < T(0));
        delta = tmin(delta, fabs(a));
    }
    const bool all_neg = !any_pos, all_pos = !any_neg;
    if (!(all_neg || all_pos)) delta = T(0);
    const T E = exp(-delta);
    // ---- pass 2: the rescaled sums (csrc/gf.cuh mix_eval without the softplus-threshold bookkeeping) ----
    T Sc = 0, Ss = 0, Sp = 0, Sd = 0;
    JF_BWD_UNROLL
    for (int k = 0; k < K; ++k) {
        const T iw = si[(size_t)k * nt], n = sn[(size_t)k * nt] * invG;
        const T a = (v - sm[(size_t)k * nt]) * iw;
        const T u = exp_neg(delta - fabs(a)), e = u * E;
        const T rx = rcp_1to2(T(1) + e);
        const T big = n * rx, small = big * u;
        if (a >= T(0)) { Sc += big; Ss += small; } else { Sc += small; Ss += big; }
        const T pt = small * iw * rx;                 // n sigma (1-sigma) / w, rescaled by 1/E
        Sp += pt;
        Sd += (a >= T(0) ? -pt : pt) * iw * (T(1) - e) * rx;
    }
    // true values: C = Sc*e^-dc, S = Ss*e^-ds, p = Sp*E, p' = Sd*E  (dc = delta if all_neg, ds = delta if all_pos)
    const T fC = all_neg ? T(1) : E, fS = all_pos ? T(1) : E;     // what is left of E after dividing by C resp. S
    T y_coefC, y_coefS;     // coefficients of (dC/dtheta)/E and (dS/dtheta)/E
    T nC, nS;               // coefficients of sigma_k and (1 - sigma_k) in their side-dependent rescaled form
    T nC_true = 0;          // coefficient of the TRUE sigma_k (terms that depend on C itself, not on log C)
    if (c.inv_type == JF_INV_ISIGMOID) {
        y_coefC = (gy - gl) * fC / Sc;
        y_coefS = -(gy + gl) * fS / Ss;
        nC = (gy - gl) / Sc;
        nS = -(gy + gl) / Ss;
    } else {
        const T Ct = Sc * (all_neg ? E : T(1)), St = Ss * (all_pos ? E : T(1));
        const T eps = T(0.5e-7);
        if (c.inv_type != JF_INV_PARTLY_PRECISE) { vbar = T(0); return false; }
        const bool upper = !(St > eps), lower = !(Ct > eps);
        if (!upper && !lower) {
            // bulk: y = Phi^-1(C), l = log sqrt(2 pi) + y^2/2 + log p   (gaussianization_flow.py:499-500, :604-606)
            const T er = (Ct <= T(0.5)) ? -erfcinv(T(2) * Ct) : erfcinv(T(2) * St);
            const T y = T(1.4142135623730951) * er;
            const T D = T(2.5066282746310002) * exp(er * er);      // 1/phi(y)
            nC_true = (gy + gl * y) * D;                            // d/dC of gy*y + gl*l
            y_coefC = nC_true * E;
            y_coefS = T(0);
            nC = T(0);
            nS = T(0);
        } else {
            // Pade tails (gaussianization_flow.py:507-536, :608-636): with L = log C + log S + log 4, F = c + L/2,
            // F2 = sqrt(F^2 - L/a):  y = +-sqrt(2 (F2 - F)),
            // l = log(F2 - F + 1/a) - log sqrt(8) - 1/2 log(F2 - F) - log F2 - log S - log C + log|1 - 2C| + log p
            const T pa = T(0.147), pc = T(2.0 / (kPi * 0.147));
            const T Lf = (log(Sc) - (all_neg ? delta : T(0))) + (log(Ss) - (all_pos ? delta : T(0))) + T(1.3862943611198906);
            const T F = pc + Lf * T(0.5);
            const T F2 = sqrt(F * F - Lf / pa);
            const T dF2 = (F - T(1) / pa) / (T(2) * F2);           // dF2/dL ; dF/dL = 1/2
            const T diff = F2 - F;
            const T yv = (upper ? T(1) : T(-1)) * sqrt(tmax(T(0), T(2) * diff));
            const T dy = (dF2 - T(0.5)) / yv;
            const T dl = (dF2 - T(0.5)) / (diff + T(1) / pa) - T(0.5) * (dF2 - T(0.5)) / diff - dF2 / F2;
            const T gL = gy * dy + gl * dl;
            nC_true = -T(2) * gl / (T(1) - T(2) * Ct);             // d/dC of gl*log|1 - 2C|
            y_coefC = (gL - gl) * fC / Sc + nC_true * E;
            y_coefS = (gL - gl) * fS / Ss;
            nC = (gL - gl) / Sc;
            nS = (gL - gl) / Ss;
        }
    }
    const T cCS = y_coefC - y_coefS;                                // dC/dtheta = -dS/dtheta for v, m, 1/w
    const T glp = gl / Sp;
    vbar = cCS * Sp + glp * Sd;
    // ---- pass 3: per-kernel gradients ----
    T nbar_dot = 0;                                                 // sum_k nbar_k n_k (normalisation Jacobian)
    const int d = c.d;
    T* gm = gp + (int64_t)(c.raw_m() + j) * sj;
    T* gw = gp + (int64_t)(c.raw_w() + j) * sj;
    T* gn = gp + (int64_t)(c.raw_n() + j) * sj;
    const int64_t step = (int64_t)d * sj;
    JF_BWD_UNROLL
    for (int k = 0; k < K; ++k) {
        const T m = sm[(size_t)k * nt], iw = si[(size_t)k * nt], g = sn[(size_t)k * nt], n = g * invG;
        const T a = (v - m) * iw;
        const T u = exp_neg(delta - fabs(a)), e = u * E;
        const T rx = rcp_1to2(T(1) + e);
        const T t = u * rx * rx;                                    // sigma (1-sigma) / E
        const T om2s = (a >= T(0) ? -(T(1) - e) : (T(1) - e)) * rx; // 1 - 2 sigma
        const T mbar = n * (-cCS * t * iw - glp * t * om2s * iw * iw);
        const T iwbar = n * (cCS * t * (v - m) + glp * t * (T(1) + om2s * a));
        // sigma_k and 1 - sigma_k: rescaled like the sums they are divided by (sigC, sigS) and true (sig_t)
        const T sigC = (a >= T(0)) ? rx : u * rx, sigS = (a >= T(0)) ? u * rx : rx;
        const T sig_t = (a >= T(0)) ? rx : u * rx * E;
        const T nbar = nC * sigC + nS * sigS + nC_true * sig_t + glp * t * iw;
        nbar_dot = fma(nbar, n, nbar_dot);
        gm[k * step] = mbar;
        // 1/w = q/(w_min q + 1), q = 1/w_max + exp(-raw): d(1/w)/d raw = -(q - 1/w_max) (1 - w_min/w)^2
        const T om = T(1) - c.w_min * iw;
        const T q = iw / om;
        gw[k * step] = -iwbar * (q - c.inv_w_max) * om * om;
        if (c.norm_mode != JF_NORM_NONE) gn[k * step] = nbar;       // finished below (needs nbar_dot)
    }
    if (c.norm_mode != JF_NORM_NONE) {
        JF_BWD_UNROLL
        for (int k = 0; k < K; ++k) {
            const T g = sn[(size_t)k * nt];
            const T gbar = (gn[k * step] - nbar_dot) * invG;        // n_k = g_k / G
            T dg;
            if (c.norm_mode == JF_NORM_REGULATED) { const T s = (g - c.n_min) / c.n_max; dg = c.n_max * s * (T(1) - s); }
            else dg = g;
            gn[k * step] = gbar * dg;
        }
    }
    return true;
}

}  // namespace jf
