// Forward AND backward of the Euclidean "g"-chain log_pdf in one kernel (training, BASELINE configs[4]).
//
//   reference: the forward is gaussianization_flow.py:995-1057 per layer (main/default.py:998-1029 layer loop), the
//   gradients are what autograd produces through :699-861 / :389-454 / :457-471; here they are closed form.
//
// Work decomposition: worker = (row, dimension j).  A warp holds 32 CONSECUTIVE rows of one dimension (the param-major
// [P, rows] block is read and its gradient written fully coalesced), the d warps of a row group meet twice per layer and
// direction through a shared-memory exchange for the Householder rotation, which every worker then applies to the whole
// row vector redundantly (d^2 multiply-adds against ~250 instructions x K for the mixture of its own dimension) -- no
// reduction trees, no shuffles.  Compared with one thread per row (csrc/gf_bwd.cuh's first kernel: 120 registers, 2.5 KB
// of local arrays, 16 warps per SM) this keeps every per-row quantity a scalar in registers and puts d x more warps on
// an SM.
//
//   forward  (l = L-1 .. 0):  u = x - offset_l ; v = Q_l^T u ; (y_j, l_j) = stage_l(v_j) ; x = y      v_j of every layer is kept
//   outputs:  base z = x, logdet = sum l_j, log N(z)          (what jf_subpdf_apply returns in the LOGPDF direction)
//   backward (l = 0 .. L-1):  zbar_j = -z_j g, lbar = g;  per dimension the closed-form mixture / inverse-CDF-stage /
//            regulator derivatives of csrc/gf_bwd.cuh (gf_elem_backward), then the reflections are undone one by one
//            (each is its own inverse) to get the gradients of the reflection vectors; offset gradient = -input gradient.
//            The gradient with respect to x falls out at the end (grad_x, optional).
#pragma once
#include "gf_bwd.cuh"
#include "gf_fb_launch.cuh"

namespace jf {

// ask L2 for the parameters this worker reads in ANOTHER layer (its own column: offset, reflection components, the K
// triples) while the current layer is being computed: the [P, rows] block is streamed from HBM exactly once per
// direction, and with ~30 warps per SM a first touch that goes to HBM is not hidden (ncu: long_scoreboard 45 % of the samples)
template <typename T>
JF_DEVINL void fb_prefetch_layer(const GfLayerC<T>& c, int D, int j, const T* prow, int64_t sj) {
    if (c.has_offset) asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + (int64_t)(c.raw_off + j) * sj));
    // the reflection components and the m / w / n blocks follow each other, every D-th parameter is this worker's
    const T* q = prow + (int64_t)(c.raw_hh() + j) * sj;
    const int64_t step = (int64_t)D * sj;
    const int nq = c.hh_iter + (c.norm_mode != JF_NORM_NONE ? 3 * c.K : 2 * c.K);
    // the 32 rows of a warp share one 128-byte line per parameter (fp32): LANE k asks for parameter k, so that one
    // instruction per warp covers 32 parameters (40 per-thread prefetches per layer were 12 % of the executed instructions)
#pragma unroll 1
    for (int k = (int)(threadIdx.x & 31); k < nq; k += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (int64_t)k * step));
}

// ---------------------------------------------------------------------------------------------------------------------
// Register-resident mixture of one (layer, dimension) for a compile-time K: the K raw triples are loaded back to back
// (3 K independent loads in flight), regulated into registers and consumed there -- no shared-memory slots, no address
// arithmetic in the loops, every loop unrolled.  The first version drove csrc/gf.cuh's slot-based helpers (rolled loops,
// 64-bit slot indices): ncu counted 8 600 instructions per (row, layer, dimension) forward + backward, 760 per mixture
// kernel, against ~200 of arithmetic.  Same formulas as mix_eval / gf_elem_backward (which stay the path for other K).
// ---------------------------------------------------------------------------------------------------------------------
#ifndef JF_FB_ROT1
#define JF_FB_ROT1 1      // 1: one worker of the row applies the reflections (0: every worker redundantly, +8 % time)
#endif
constexpr int kFbFastK = 10;   // the default num_kde of the reference (flow_options.py): register-resident path

template <typename T, int K>
struct FbMix {
    T m[K], iw[K], g[K];   // means, regulated 1/width, unnormalised regulated weights
    T G;                   // sum of the weights
};

// pm: address of (raw_m + j) of this row; the m / w / n blocks follow each other, K * d parameters each (step = d * sj)
template <typename T, int K>
JF_DEVINL void fb_load(const GfLayerC<T>& c, const T* pm, int64_t step, FbMix<T, K>& M) {
    T rw[K], rn[K];
    const T* q = pm;                       // (a walking pointer: k * step as a 64-bit product costs 5 instructions per load)
#pragma unroll
    for (int k = 0; k < K; ++k) { M.m[k] = *q; q += step; }
#pragma unroll
    for (int k = 0; k < K; ++k) { rw[k] = *q; q += step; }
    if (c.norm_mode != JF_NORM_NONE) {
#pragma unroll
        for (int k = 0; k < K; ++k) { rn[k] = *q; q += step; }
    }
    T G = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) M.iw[k] = regulate_inv_width(rw[k], c.w_min, c.inv_w_max);
    if (c.norm_mode == JF_NORM_REGULATED) {
#pragma unroll
        for (int k = 0; k < K; ++k) { M.g[k] = regulate_norm(rn[k], c.n_min, c.n_max); G += M.g[k]; }
    } else if (c.norm_mode == JF_NORM_RAW) {
        T nmax = -Num<T>::big;
#pragma unroll
        for (int k = 0; k < K; ++k) nmax = tmax(nmax, rn[k]);
#pragma unroll
        for (int k = 0; k < K; ++k) { M.g[k] = exp(rn[k] - nmax); G += M.g[k]; }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) M.g[k] = T(1);
        G = T(K);
    }
    M.G = G;
}

// rescaling exponent (see mix_eval): 0 unless every kernel lies on one side of x
template <typename T, int K>
JF_DEVINL T fb_delta(const FbMix<T, K>& M, T x, bool& all_neg, bool& all_pos) {
    bool any_pos = false, any_neg = false;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const bool pos = x >= M.m[k];                  // 1/w > 0: sign(a_k) = sign(x - m_k)
        any_pos = any_pos || pos;
        any_neg = any_neg || !pos;
    }
    all_neg = !any_pos;
    all_pos = !any_neg;
    T delta = 0;
    if (all_neg || all_pos) {
        delta = Num<T>::big;
#pragma unroll
        for (int k = 0; k < K; ++k) delta = tmin(delta, fabs((x - M.m[k]) * M.iw[k]));
    }
    return delta;
}

// forward of one element: load + regulate + mixture + inverse-CDF stage
template <typename T, int K>
__device__ __noinline__ void fb_fwd_elem(const GfLayerC<T>& c, const T* pm, int64_t step, T x, T& y, T& logd) {
    FbMix<T, K> M;
    fb_load<T, K>(c, pm, step, M);
    bool all_neg, all_pos;
    const T delta = fb_delta<T, K>(M, x, all_neg, all_pos);
    const T E = (delta > T(0)) ? exp_neg(-delta) : T(1);
    T big_p = 0, small_p = 0, big_n = 0, small_n = 0, Sp = 0, amin = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const T iw = M.iw[k];
        const T a = (x - M.m[k]) * iw;
        const T u = exp_neg(delta - fabs(a));
        const T rx = rcp_1to2(fma(u, E, T(1)));        // exact sigma(|a|)
        const T nr = M.g[k] * rx, nur = nr * u;
        const bool pos = a >= T(0);
        big_p += pos ? nr : T(0);
        small_p += pos ? nur : T(0);
        big_n += pos ? T(0) : nr;
        small_n += pos ? T(0) : nur;
        Sp = fma(nur * iw, rx, Sp);
        amin = tmin(amin, a);
    }
    T ex = 0, qc = 0;
    if (amin < T(-20)) {                               // softplus-threshold quirk of the reference, see mix_eval (rare)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const T iw = M.iw[k];
            const T a = (x - M.m[k]) * iw;
            if (a < T(-20)) {
                const T u = exp_neg(delta - fabs(a));
                const T e = u * E;
                const T rx = rcp_1to2(T(1) + e);
                const T nq = M.g[k] * e * rx;
                ex += nq;
                qc = fma(nq, u, qc);
                Sp = fma(nq * u * iw, T(1) + rx, Sp);
            }
        }
    }
    const T invG = rcp_pos_(M.G);
    MixVal<T> v;
    v.Sc = (big_p + small_n + qc) * invG;
    v.Ss = (small_p + big_n) * invG;
    v.Sp = Sp * invG;
    v.ex = ex * invG;
    v.Sd = 0;
    v.E = E;
    v.dc = all_neg ? delta : T(0);
    v.ds = all_pos ? delta : T(0);
    v.dp = delta;
    inv_stage(c.inv_type, v, y, logd);
}

// coefficients of the inverse-CDF stage's reverse pass (linear in the upstream gradients gy of y and gl of l = log y'):
// see csrc/gf_bwd.cuh gf_elem_backward for the derivation
template <typename T>
struct FbCoef {
    T y_coefC, y_coefS;     // coefficients of (dC/dtheta)/E and (dS/dtheta)/E
    T nC, nS;               // coefficients of sigma_k and (1 - sigma_k) in their side-dependent rescaled form
    T nC_true;              // coefficient of the TRUE sigma_k
    bool ok;
};
template <typename T>
JF_DEVINL FbCoef<T> fb_coef(const GfLayerC<T>& c, T Sc, T Ss, T E, T delta, bool all_neg, bool all_pos, T gy, T gl) {
    FbCoef<T> r;
    r.ok = true;
    r.nC_true = 0;
    const T fC = all_neg ? T(1) : E, fS = all_pos ? T(1) : E;
    if (c.inv_type == JF_INV_ISIGMOID) {
        r.y_coefC = (gy - gl) * fC / Sc;
        r.y_coefS = -(gy + gl) * fS / Ss;
        r.nC = (gy - gl) / Sc;
        r.nS = -(gy + gl) / Ss;
        return r;
    }
    const T Ct = Sc * (all_neg ? E : T(1)), St = Ss * (all_pos ? E : T(1));
    const T eps = T(0.5e-7);
    if (c.inv_type != JF_INV_PARTLY_PRECISE) { r.ok = false; r.y_coefC = r.y_coefS = r.nC = r.nS = 0; return r; }
    const bool upper = !(St > eps), lower = !(Ct > eps);
    if (!upper && !lower) {
        const T er = (Ct <= T(0.5)) ? -erfcinv(T(2) * Ct) : erfcinv(T(2) * St);
        const T y = T(1.4142135623730951) * er;
        const T Dn = T(2.5066282746310002) * exp(er * er);      // 1/phi(y)
        r.nC_true = (gy + gl * y) * Dn;
        r.y_coefC = r.nC_true * E;
        r.y_coefS = T(0);
        r.nC = T(0);
        r.nS = T(0);
    } else {
        const T pa = T(0.147), pc = T(2.0 / (kPi * 0.147));
        const T Lf = (log(Sc) - (all_neg ? delta : T(0))) + (log(Ss) - (all_pos ? delta : T(0))) + T(1.3862943611198906);
        const T F = pc + Lf * T(0.5);
        const T F2 = sqrt(F * F - Lf / pa);
        const T dF2 = (F - T(1) / pa) / (T(2) * F2);
        const T diff = F2 - F;
        const T yv = (upper ? T(1) : T(-1)) * sqrt(tmax(T(0), T(2) * diff));
        const T dy = (dF2 - T(0.5)) / yv;
        const T dl = (dF2 - T(0.5)) / (diff + T(1) / pa) - T(0.5) * (dF2 - T(0.5)) / diff - dF2 / F2;
        const T gL = gy * dy + gl * dl;
        r.nC_true = -T(2) * gl / (T(1) - T(2) * Ct);
        r.y_coefC = (gL - gl) * fC / Sc + r.nC_true * E;
        r.y_coefS = (gL - gl) * fS / Ss;
        r.nC = (gL - gl) / Sc;
        r.nS = (gL - gl) / Ss;
    }
    return r;
}

// backward of one element (v -> (y, l); gy, gl upstream): writes the raw-parameter gradients of this (layer, dimension)
// through gm (address of (raw_m + j) in the gradient block), returns d/dv in vbar; false for rows outside the supported
// stages.  Formulas: csrc/gf_bwd.cuh gf_elem_backward.
// SDIR: the element is traversed in the SAMPLING direction (y -> v): gy is then the cotangent of v, gl that of log_pdf, and
// vbar returns the cotangent of y.
template <typename T, int K, bool SDIR = false>
__device__ __noinline__ bool fb_bwd_elem(const GfLayerC<T>& c, const T* pm, T* gm, int64_t step, T v, T gy, T gl, T& vbar) {
    FbMix<T, K> M;
    fb_load<T, K>(c, pm, step, M);
    const T invG = rcp_pos_(M.G);
    bool all_neg, all_pos;
    const T delta = fb_delta<T, K>(M, v, all_neg, all_pos);
    const T E = (delta > T(0)) ? exp_neg(-delta) : T(1);
    T Sc = 0, Ss = 0, Sp = 0, Sd = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const T iw = M.iw[k], n = M.g[k] * invG;
        const T a = (v - M.m[k]) * iw;
        const T u = exp_neg(delta - fabs(a)), e = u * E;
        const T rx = rcp_1to2(T(1) + e);
        const T big = n * rx, small = big * u;
        const bool pos = a >= T(0);
        Sc += pos ? big : small;
        Ss += pos ? small : big;
        const T pt = small * iw * rx;                  // n sigma (1-sigma) / w, rescaled by 1/E
        Sp += pt;
        const T dd = pt * iw * (T(1) - e) * rx;
        Sd += pos ? -dd : dd;
    }
    FbCoef<T> cf;
    if (SDIR) {
        // sampling direction: y -> v = f^-1(y), log_pdf gains l(v).  With A = f'(v) and l_v (the coefficients are linear in
        // (gy, gl)): cotangent of y is w = (vbar_in + gl l_v) / A, the parameter gradients are those of the log_pdf
        // direction for (gy, gl) = (-w, gl)
        const FbCoef<T> ca = fb_coef<T>(c, Sc, Ss, E, delta, all_neg, all_pos, T(1), T(0));
        const FbCoef<T> cb = fb_coef<T>(c, Sc, Ss, E, delta, all_neg, all_pos, T(0), T(1));
        if (!ca.ok) { vbar = T(0); return false; }
        const T A = (ca.y_coefC - ca.y_coefS) * Sp;
        const T Bv = (cb.y_coefC - cb.y_coefS) * Sp + Sd / Sp;
        const T w = (gy + gl * Bv) / A;
        cf = fb_coef<T>(c, Sc, Ss, E, delta, all_neg, all_pos, -w, gl);
        vbar = w;
    } else {
        cf = fb_coef<T>(c, Sc, Ss, E, delta, all_neg, all_pos, gy, gl);
        if (!cf.ok) { vbar = T(0); return false; }
    }
    const T nC = cf.nC, nS = cf.nS, nC_true = cf.nC_true;
    const T cCS = cf.y_coefC - cf.y_coefS;
    const T glp = gl / Sp;
    if (!SDIR) vbar = cCS * Sp + glp * Sd;
    T nb[K];
    T nbar_dot = 0;
    T* gq_m = gm;
    T* gq_w = gm + (int64_t)K * step;
    T* gq_n = gq_w + (int64_t)K * step;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const T m = M.m[k], iw = M.iw[k], n = M.g[k] * invG;
        const T a = (v - m) * iw;
        const T u = exp_neg(delta - fabs(a)), e = u * E;
        const T rx = rcp_1to2(T(1) + e);
        const T t = u * rx * rx;                                    // sigma (1-sigma) / E
        const bool pos = a >= T(0);
        const T om2s = (pos ? -(T(1) - e) : (T(1) - e)) * rx;       // 1 - 2 sigma
        const T mbar = n * (-cCS * t * iw - glp * t * om2s * iw * iw);
        const T iwbar = n * (cCS * t * (v - m) + glp * t * (T(1) + om2s * a));
        const T ur = u * rx;
        const T sigC = pos ? rx : ur, sigS = pos ? ur : rx;
        const T sig_t = pos ? rx : ur * E;
        nb[k] = nC * sigC + nS * sigS + nC_true * sig_t + glp * t * iw;
        nbar_dot = fma(nb[k], n, nbar_dot);
        *gq_m = mbar;
        gq_m += step;
        // 1/w = q/(w_min q + 1), q = 1/w_max + exp(-raw): d(1/w)/d raw = -(q - 1/w_max) (1 - w_min/w)^2
        const T om = T(1) - c.w_min * iw;
        const T q = iw * rcp_pos_(om);
        *gq_w = -iwbar * (q - c.inv_w_max) * om * om;
        gq_w += step;
    }
    if (c.norm_mode != JF_NORM_NONE) {
        const T inv_nmax = c.norm_mode == JF_NORM_REGULATED ? T(1) / c.n_max : T(0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const T g = M.g[k];
            const T gbar = (nb[k] - nbar_dot) * invG;               // n_k = g_k / G
            T dg = g;
            if (c.norm_mode == JF_NORM_REGULATED) { const T s = (g - c.n_min) * inv_nmax; dg = c.n_max * s * (T(1) - s); }
            *gq_n = gbar * dg;
            gq_n += step;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// "t" (affine) layers inside the chain: x = L z + mu with a lower-triangular L (csrc/gf.cuh mvn_layer_*; reference
// layers/euclidean/multivariate_normal.py:228-272).  A "t" layer costs ~d^2 operations per row against ~2500 x d for a "g"
// layer, so ONE worker of the row (dimension 0) does the whole vector: the others hand their component over through the
// exchange and wait.
//   log_pdf direction: y = x - mu, L z = y, logdet -= sum log w_i.  Reverse pass for cotangents zb of z and gl of logdet:
//   yb = L^-T zb, Lbar_ij = -yb_i z_j, wbar_i = -yb_i z_i - gl / w_i, mubar = -yb, xbar = yb.
//   sampling direction (MODE 1): x = L z + mu: zb = L^T xb, Lbar_ij = xb_i z_j, wbar_i = xb_i z_i - gl / w_i, mubar = xb.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T> JF_DEVINL T fb_dw_draw(T w, T w_min, T inv_w_max) {
    // w = w_min + 1/q, q = 1/w_max + exp(-raw):  dw/draw = (q - 1/w_max) / q^2
    const T r = w - w_min;                              // 1/q
    return (T(1) / r - inv_w_max) * r * r;
}

template <typename T, int D, bool SDIR>
__device__ __noinline__ void fb_mvn_backward(const GfLayerC<T>& c, const T* prow, T* grow, int64_t sj, const T* z, T* zb, T gl) {
    const int cov = c.inv_type;
    const T* q = prow + (int64_t)(c.raw_off + (c.has_offset ? D : 0)) * sj;
    T* gq = grow + (int64_t)(c.raw_off + (c.has_offset ? D : 0)) * sj;
    if (SDIR && c.has_offset) {
        for (int i = 0; i < D; ++i) grow[(int64_t)(c.raw_off + i) * sj] = zb[i];
    }
    if (cov == 1) {
        T w, iw;
        regulate_width(q[0], c.w_min, c.inv_w_max, w, iw);
        T s = 0;
        for (int i = 0; i < D; ++i) s = fma(zb[i], z[i], s);
        // log_pdf: z = y / w (dz/dw = -z/w);  sampling: x = w z (dx/dw = z)
        const T wbar = (SDIR ? s : -s * iw) - gl * T(D) * iw;
        gq[0] = wbar * fb_dw_draw(w, c.w_min, c.inv_w_max);
        for (int i = 0; i < D; ++i) zb[i] *= SDIR ? w : iw;
    } else if (cov >= 2) {
        T w[D], iw[D], out[D];
        for (int i = 0; i < D; ++i) regulate_width(q[(int64_t)i * sj], c.w_min, c.inv_w_max, w[i], iw[i]);
        if (SDIR) {
            for (int i = 0; i < D; ++i) {                   // zb_out = L^T xb
                T acc = w[i] * zb[i];
                if (cov == 3)
                    for (int k = i + 1; k < D; ++k) acc = fma(q[(int64_t)(D + mvn_lower_index(D, k, i)) * sj], zb[k], acc);
                out[i] = acc;
            }
            for (int i = 0; i < D; ++i) {
                gq[(int64_t)i * sj] = (zb[i] * z[i] - gl * iw[i]) * fb_dw_draw(w[i], c.w_min, c.inv_w_max);
                if (cov == 3)
                    for (int jj = 0; jj < i; ++jj) gq[(int64_t)(D + mvn_lower_index(D, i, jj)) * sj] = zb[i] * z[jj];
            }
        } else {
            for (int ii = 0; ii < D; ++ii) {                // yb = L^-T zb (back substitution)
                const int i = D - 1 - ii;
                T acc = zb[i];
                if (cov == 3)
                    for (int k = i + 1; k < D; ++k) acc = fma(-q[(int64_t)(D + mvn_lower_index(D, k, i)) * sj], out[k], acc);
                out[i] = acc * iw[i];
            }
            for (int i = 0; i < D; ++i) {
                gq[(int64_t)i * sj] = (-out[i] * z[i] - gl * iw[i]) * fb_dw_draw(w[i], c.w_min, c.inv_w_max);
                if (cov == 3)
                    for (int jj = 0; jj < i; ++jj) gq[(int64_t)(D + mvn_lower_index(D, i, jj)) * sj] = -out[i] * z[jj];
            }
        }
        for (int i = 0; i < D; ++i) zb[i] = out[i];
    }
    if (!SDIR && c.has_offset) {
        for (int i = 0; i < D; ++i) grow[(int64_t)(c.raw_off + i) * sj] = -zb[i];
    }
}

template <typename T, int D>
__device__ __noinline__ void fb_mvn_forward(const GfLayerC<T>& c, const T* prow, int64_t sj, T* x, T& ld) {
    T v[D];
    for (int i = 0; i < D; ++i) v[i] = x[i];
    mvn_layer_logpdf<T, D>(v, ld, c, D, prow, sj);
    for (int i = 0; i < D; ++i) x[i] = v[i];
}

JF_DEVINL void fb_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// MODE 2: log_pdf forward only (what jf_subpdf_apply runs for per-row parameters: same workers, same register-resident
// mixture, no reverse pass).  MODE 0: log_pdf forward + backward.  MODE 1: backward of the SAMPLING direction at the sample x = T(z; theta) (a.in): the
// layers' inputs are recovered by running the closed-form log_pdf direction from x (no root finder), then the chain is
// walked from the x side to the z side with the implicit-function form of every element (fb_bwd_elem<SDIR>): cotangents
// grad_out_x of x and grad_logp of log_pdf(x) in, parameter gradients and (optionally, grad_x) the cotangent of z out.
template <typename T, int D, int MODE>
__global__ void __launch_bounds__(fb_threads(D), fb_min_blocks(D, sizeof(T))) gf_chain_fb_kernel(const __grid_constant__ GfFbArgs<T> g) {
    constexpr int NG = fb_groups(D);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* slots = reinterpret_cast<T*>(smem_raw);
    const SubPdfArgs<T>& a = g.a;
    const int L = a.n_layers;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rg = warp / D, j = warp - rg * D;
    // exchange of this row group: field f of lane at ex[f * 32]; fields: X[D] | XB[D] | H[hh][D]
    T* ex = slots + (size_t)3 * g.kmax * fb_threads(D) + (size_t)rg * (g.hh_max * D + 2 * D) * 32 + lane;
    constexpr int fX = 0, fXB = D, fH = 2 * D;
    const int bar_id = 1 + rg;
    const int64_t row_raw = ((int64_t)blockIdx.x * NG + rg) * 32 + lane;
    const bool live = row_raw < a.B;
    const int64_t row = live ? row_raw : a.B - 1;      // (dead lanes load valid memory, store nothing)
    const T* prow = a.params + row * a.sr;
    T* grow = g.grad_params + row * a.sr;
    const int64_t sj = a.sj;

    T xj = a.in[row * a.ld_in + j];
    if (MODE == 2 && a.emb_out != nullptr && live) a.emb_out[row * a.ld_emb + j] = xj;      // Euclidean: the target itself
    T ld_acc = 0;
    T vsave[JF_MAX_LAYERS];                            // pre-stage value of this dimension in every layer
    // ---- forward ----
#pragma unroll 1
    for (int l = L - 1; l >= 0; --l) {
        const GfLayerC<T>& c = g.layers[l];
        if (c.kind == 1) {
            // "t" layer: worker 0 of the row maps the whole vector (see fb_mvn_backward); vsave keeps the layer's OUTPUT
            ex[(fX + j) * 32] = xj;
            fb_bar(bar_id, 32 * D);
            if (j == 0 && live) {
                T X[D], ldt = 0;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) X[jj] = ex[(fX + jj) * 32];
                fb_mvn_forward<T, D>(c, prow, sj, X, ldt);
                ld_acc += ldt;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) ex[(fXB + jj) * 32] = X[jj];
            }
            fb_bar(bar_id, 32 * D);
            xj = ex[(fXB + j) * 32];
            vsave[l] = xj;
            fb_bar(bar_id, 32 * D);
            continue;
        }
        if (l > 0 && g.layers[l - 1].kind == 0) fb_prefetch_layer<T>(g.layers[l - 1], D, j, prow, sj);
        if (c.has_offset) xj -= prow[(int64_t)(c.raw_off + j) * sj];
        if (c.hh_iter > 0) {
            ex[(fX + j) * 32] = xj;
            {
                const T* ph = prow + (int64_t)(c.raw_hh() + j) * sj;
                T* eh = ex + (fH + j) * 32;
#pragma unroll 5
                for (int i = 0; i < c.hh_iter; ++i) { *eh = *ph; ph += (int64_t)D * sj; eh += D * 32; }
            }
            fb_bar(bar_id, 32 * D);
#if JF_FB_ROT1
            if (MODE != 1) {
                // experiment: ONE worker of the row applies the reflections (no redundant work), the others wait
                if (j == 0) {
                    T X[D];
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) X[jj] = ex[(fX + jj) * 32];
#pragma unroll 1
                    for (int i = 0; i < c.hh_iter; ++i) {
                        T w[D], dot = 0, nrm = 0;
#pragma unroll
                        for (int jj = 0; jj < D; ++jj) {
                            w[jj] = ex[(fH + i * D + jj) * 32];
                            dot = fma(w[jj], X[jj], dot);
                            nrm = fma(w[jj], w[jj], nrm);
                        }
                        const T cc = T(2) * dot * rcp_pos_(nrm);
#pragma unroll
                        for (int jj = 0; jj < D; ++jj) X[jj] = fma(-cc, w[jj], X[jj]);
                    }
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) ex[(fXB + jj) * 32] = X[jj];
                }
                fb_bar(bar_id, 32 * D);
                xj = ex[(fXB + j) * 32];
            } else
#endif
            {
            T X[D];
#pragma unroll
            for (int jj = 0; jj < D; ++jj) X[jj] = ex[(fX + jj) * 32];
#pragma unroll 1
            for (int i = 0; i < c.hh_iter; ++i) {
                T w[D], dot = 0, nrm = 0;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) {
                    w[jj] = ex[(fH + i * D + jj) * 32];
                    dot = fma(w[jj], X[jj], dot);
                    nrm = fma(w[jj], w[jj], nrm);
                }
                const T cc = T(2) * dot * rcp_pos_(nrm);
#pragma unroll
                for (int jj = 0; jj < D; ++jj) X[jj] = fma(-cc, w[jj], X[jj]);
            }
#pragma unroll
            for (int jj = 0; jj < D; ++jj) xj = (jj == j) ? X[jj] : xj;
            fb_bar(bar_id, 32 * D);            // everybody has read the exchange
            }
        }
        vsave[l] = xj;
        if (live) {
            T y, logd;
            if (c.K == kFbFastK) {
                fb_fwd_elem<T, kFbFastK>(c, prow + (int64_t)(c.raw_m() + j) * sj, (int64_t)D * sj, xj, y, logd);
            } else {
                const MixView<T> mv = regulate_to_slots<T>(c, c.K, j, prow, sj, slots, false);
                gf_eval_logpdf<T>(mv, c.inv_type, xj, y, logd);
            }
            xj = y;
            ld_acc += logd;
        }
    }
    if (MODE == 0 || MODE == 2) {
        // ---- forward outputs: base point, logdet, log N(z) ----
        if (live && a.out != nullptr) a.out[row * a.ld_out + j] = xj;
        if (MODE == 2 || a.logdet_out != nullptr || a.logbase_out != nullptr) {
            ex[(fX + j) * 32] = ld_acc;
            ex[(fXB + j) * 32] = xj * xj;
            fb_bar(bar_id, 32 * D);
            if (j == 0 && live) {
                T ld = a.logdet_in != nullptr ? a.logdet_in[row] : T(0), zsq = 0;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) { ld += ex[(fX + jj) * 32]; zsq += ex[(fXB + jj) * 32]; }
                // (a non-finite base coordinate makes zsq non-finite: one count per row)
                if (!finite_(ld) || (MODE == 2 && !finite_(zsq))) status_add(a.status, JF_STATUS_NONFINITE, 1);
                if (a.logdet_out != nullptr) a.logdet_out[row] = ld;
                if (a.logbase_out != nullptr)
                    a.logbase_out[row] = (a.logbase_in != nullptr ? a.logbase_in[row] : T(0)) - T(0.5) * zsq - T(D) * T(kLogSqrt2Pi);
            }
            fb_bar(bar_id, 32 * D);
        }
        if (MODE == 2) return;                         // forward only: the (row, dimension) log_pdf kernel of jf_subpdf_apply
    }
    if (MODE == 1) {
        // ---- backward of the sampling direction: x side first ----
        T xb = g.grad_out_x ? g.grad_out_x[row * g.ld_go + j] : T(0);
        const T gl = g.grad_logp ? g.grad_logp[row] : T(0);
        int n_bad = 0;
#pragma unroll 1
        for (int l = L - 1; l >= 0; --l) {
            const GfLayerC<T>& c = g.layers[l];
            const T v = vsave[l];
            if (c.kind == 1) {
                ex[(fX + j) * 32] = v;
                ex[(fXB + j) * 32] = xb;
                fb_bar(bar_id, 32 * D);
                if (j == 0 && live) {
                    T Z[D], ZB[D];
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) { Z[jj] = ex[(fX + jj) * 32]; ZB[jj] = ex[(fXB + jj) * 32]; }
                    fb_mvn_backward<T, D, true>(c, prow, grow, sj, Z, ZB, gl);
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) ex[(fX + jj) * 32] = ZB[jj];
                }
                fb_bar(bar_id, 32 * D);
                xb = ex[(fX + j) * 32];
                fb_bar(bar_id, 32 * D);
                continue;
            }
            if (l > 0 && g.layers[l - 1].kind == 0) fb_prefetch_layer<T>(g.layers[l - 1], D, j, prow, sj);
            if (c.has_offset && live) grow[(int64_t)(c.raw_off + j) * sj] = xb;      // x = Q v + offset
            if (c.hh_iter > 0) {
                ex[(fX + j) * 32] = v;
                ex[(fXB + j) * 32] = xb;
                {
                    const T* ph = prow + (int64_t)(c.raw_hh() + j) * sj;
                    T* eh = ex + (fH + j) * 32;
#pragma unroll 5
                    for (int i = 0; i < c.hh_iter; ++i) { *eh = *ph; ph += (int64_t)D * sj; eh += D * 32; }
                }
                fb_bar(bar_id, 32 * D);
                T V[D], XB[D];
#pragma unroll
                for (int jj = 0; jj < D; ++jj) { V[jj] = ex[(fX + jj) * 32]; XB[jj] = ex[(fXB + jj) * 32]; }
                // t_0 = H_0 ... H_{n-1} v  (the rotated vector, = x - offset)
#pragma unroll 1
                for (int i = c.hh_iter - 1; i >= 0; --i) {
                    T w[D], dot = 0, nrm = 0;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) {
                        w[jj] = ex[(fH + i * D + jj) * 32];
                        dot = fma(w[jj], V[jj], dot);
                        nrm = fma(w[jj], w[jj], nrm);
                    }
                    const T cc = T(2) * dot * rcp_pos_(nrm);
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) V[jj] = fma(-cc, w[jj], V[jj]);
                }
                T t_own = v, xb_own = xb;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) t_own = (jj == j) ? V[jj] : t_own;
                // undo the reflections from the output side: t_i = H_i t_{i+1}, cotangent of t_i known
                const int64_t hstep = (int64_t)D * sj;
                T* gh = grow + (int64_t)(c.raw_hh() + j) * sj;
#pragma unroll 1
                for (int i = 0; i < c.hh_iter; ++i) {
                    T w[D], s_ = 0, aa = 0, bb = 0;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) {
                        w[jj] = ex[(fH + i * D + jj) * 32];
                        s_ = fma(w[jj], w[jj], s_);
                        aa = fma(w[jj], V[jj], aa);         // w . t_i ( = -(w . t_{i+1}) )
                        bb = fma(w[jj], XB[jj], bb);
                    }
                    const T is = rcp_pos_(s_);
                    const T ain = -aa;
                    const T w_own = ex[(fH + i * D + j) * 32];
                    const T ca = T(2) * aa * is, cb = T(2) * bb * is;
                    const T tin = fma(-ca, w_own, t_own);
                    if (live) *gh = -T(2) * is * (bb * tin + ain * xb_own) + T(4) * ain * bb * is * is * w_own;
                    gh += hstep;
                    xb_own = fma(-cb, w_own, xb_own);
                    t_own = tin;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) { V[jj] = fma(-ca, w[jj], V[jj]); XB[jj] = fma(-cb, w[jj], XB[jj]); }
                }
                xb = xb_own;                                  // cotangent of v
                fb_bar(bar_id, 32 * D);
            }
            if (live) {
                T ybar;
                bool ok;
                if (c.K == kFbFastK) {
                    ok = fb_bwd_elem<T, kFbFastK, true>(c, prow + (int64_t)(c.raw_m() + j) * sj, grow + (int64_t)(c.raw_m() + j) * sj,
                                                        (int64_t)D * sj, v, xb, gl, ybar);
                } else {
                    const T Gs = bwd_regulate<T>(c, c.K, j, prow, sj, slots);
                    T A_, Bv_, dummy;
                    ok = gf_elem_backward<T>(c, c.K, j, v, T(1), T(0), Gs, slots, grow, sj, A_);
                    if (ok) {
                        gf_elem_backward<T>(c, c.K, j, v, T(0), T(1), Gs, slots, grow, sj, Bv_);
                        ybar = (xb + gl * Bv_) / A_;
                        gf_elem_backward<T>(c, c.K, j, v, -ybar, gl, Gs, slots, grow, sj, dummy);
                    }
                }
                if (!ok) {
                    ++n_bad;
                    ybar = T(0);
#pragma unroll 1
                    for (int k = 0; k < c.K; ++k) {
                        grow[(int64_t)(c.raw_m() + k * D + j) * sj] = T(0);
                        grow[(int64_t)(c.raw_w() + k * D + j) * sj] = T(0);
                        if (c.norm_mode != JF_NORM_NONE) grow[(int64_t)(c.raw_n() + k * D + j) * sj] = T(0);
                    }
                }
                xb = ybar;
            }
        }
        if (g.grad_x != nullptr && live) g.grad_x[row * g.ld_gx + j] = xb;
        if (n_bad) status_add(a.status, JF_STATUS_OUT_OF_RANGE, n_bad);
        return;
    }
    // ---- backward ----
    const T gr = g.grad_logp ? g.grad_logp[row] : T(1);
    T xb = -xj * gr;                                   // d/dz_j of sum_j log N(z_j)
    int n_bad = 0;
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
        const GfLayerC<T>& c = g.layers[l];
        const T v = vsave[l];
        if (c.kind == 1) {
            ex[(fX + j) * 32] = v;
            ex[(fXB + j) * 32] = xb;
            fb_bar(bar_id, 32 * D);
            if (j == 0 && live) {
                T Z[D], ZB[D];
#pragma unroll
                for (int jj = 0; jj < D; ++jj) { Z[jj] = ex[(fX + jj) * 32]; ZB[jj] = ex[(fXB + jj) * 32]; }
                fb_mvn_backward<T, D, false>(c, prow, grow, sj, Z, ZB, gr);
#pragma unroll
                for (int jj = 0; jj < D; ++jj) ex[(fX + jj) * 32] = ZB[jj];
            }
            fb_bar(bar_id, 32 * D);
            xb = ex[(fX + j) * 32];
            fb_bar(bar_id, 32 * D);
            continue;
        }
        if (l + 1 < L && g.layers[l + 1].kind == 0) fb_prefetch_layer<T>(g.layers[l + 1], D, j, prow, sj);
        if (live) {
            T vbar;
            bool ok;
            if (c.K == kFbFastK) {
                ok = fb_bwd_elem<T, kFbFastK>(c, prow + (int64_t)(c.raw_m() + j) * sj, grow + (int64_t)(c.raw_m() + j) * sj,
                                              (int64_t)D * sj, v, xb, gr, vbar);
            } else {
                const T Gs = bwd_regulate<T>(c, c.K, j, prow, sj, slots);
                ok = gf_elem_backward<T>(c, c.K, j, v, xb, gr, Gs, slots, grow, sj, vbar);
            }
            if (!ok) {
                ++n_bad;
#pragma unroll 1
                for (int k = 0; k < c.K; ++k) {        // no stage gradient for rows in the Pade tails
                    grow[(int64_t)(c.raw_m() + k * D + j) * sj] = T(0);
                    grow[(int64_t)(c.raw_w() + k * D + j) * sj] = T(0);
                    if (c.norm_mode != JF_NORM_NONE) grow[(int64_t)(c.raw_n() + k * D + j) * sj] = T(0);
                }
            }
            xb = vbar;
        }
        if (c.hh_iter > 0) {
            // v = H_{n-1} ... H_0 u: undo reflection by reflection (each one is its own inverse)
            ex[(fX + j) * 32] = v;
            ex[(fXB + j) * 32] = xb;
            {
                const T* ph = prow + (int64_t)(c.raw_hh() + j) * sj;
                T* eh = ex + (fH + j) * 32;
#pragma unroll 5
                for (int i = 0; i < c.hh_iter; ++i) { *eh = *ph; ph += (int64_t)D * sj; eh += D * 32; }
            }
            fb_bar(bar_id, 32 * D);
#if JF_FB_ROT1
            if (j == 0) {
                T V[D], XB[D];
#pragma unroll
                for (int jj = 0; jj < D; ++jj) { V[jj] = ex[(fX + jj) * 32]; XB[jj] = ex[(fXB + jj) * 32]; }
                const int64_t hstep1 = (int64_t)D * sj;
                T* gh1 = grow + (int64_t)c.raw_hh() * sj + (int64_t)(c.hh_iter - 1) * hstep1;
#pragma unroll 1
                for (int i = c.hh_iter - 1; i >= 0; --i) {
                    T w[D], s_ = 0, aa = 0, bb = 0;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) {
                        w[jj] = ex[(fH + i * D + jj) * 32];
                        s_ = fma(w[jj], w[jj], s_);
                        aa = fma(w[jj], V[jj], aa);
                        bb = fma(w[jj], XB[jj], bb);
                    }
                    const T is = rcp_pos_(s_);
                    const T ain = -aa;
                    const T ca = T(2) * aa * is, cb = T(2) * bb * is, c4 = T(4) * ain * bb * is * is, m2 = -T(2) * is;
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) {
                        const T xin = fma(-ca, w[jj], V[jj]);
                        if (live) gh1[(int64_t)jj * sj] = fma(m2, bb * xin + ain * XB[jj], c4 * w[jj]);
                        XB[jj] = fma(-cb, w[jj], XB[jj]);
                        V[jj] = xin;
                    }
                    gh1 -= hstep1;
                }
#pragma unroll
                for (int jj = 0; jj < D; ++jj) ex[(fXB + jj) * 32] = XB[jj];
            }
            fb_bar(bar_id, 32 * D);
            const T xb_own = ex[(fXB + j) * 32];
#else
            T V[D], XB[D];
#pragma unroll
            for (int jj = 0; jj < D; ++jj) { V[jj] = ex[(fX + jj) * 32]; XB[jj] = ex[(fXB + jj) * 32]; }
            T v_own = v, xb_own = xb;
            const int64_t hstep = (int64_t)D * sj;
            T* gh = grow + (int64_t)(c.raw_hh() + j) * sj + (int64_t)(c.hh_iter - 1) * hstep;
#pragma unroll 1
            for (int i = c.hh_iter - 1; i >= 0; --i) {
                T w[D], s = 0, aa = 0, bb = 0;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) {
                    w[jj] = ex[(fH + i * D + jj) * 32];
                    s = fma(w[jj], w[jj], s);
                    aa = fma(w[jj], V[jj], aa);        // w . x_out ( = -(w . x_in) )
                    bb = fma(w[jj], XB[jj], bb);
                }
                const T is = rcp_pos_(s);
                const T ain = -aa;
                const T w_own = ex[(fH + i * D + j) * 32];
                const T ca = T(2) * aa * is, cb = T(2) * bb * is;
                const T xin = fma(-ca, w_own, v_own);
                if (live) *gh = -T(2) * is * (bb * xin + ain * xb_own) + T(4) * ain * bb * is * is * w_own;
                gh -= hstep;
                xb_own = fma(-cb, w_own, xb_own);
                v_own = xin;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) { V[jj] = fma(-ca, w[jj], V[jj]); XB[jj] = fma(-cb, w[jj], XB[jj]); }
            }
#endif
            xb = xb_own;
#if !JF_FB_ROT1
            fb_bar(bar_id, 32 * D);
#endif
        }
        if (c.has_offset && live) grow[(int64_t)(c.raw_off + j) * sj] = -xb;
    }
    if (g.grad_x != nullptr && live) g.grad_x[row * g.ld_gx + j] = xb;
    if (n_bad) status_add(a.status, JF_STATUS_OUT_OF_RANGE, n_bad);
}

}  // namespace jf
