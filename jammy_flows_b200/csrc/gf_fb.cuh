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

JF_DEVINL void fb_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <typename T, int D>
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
    T ld_acc = 0;
    T vsave[JF_MAX_LAYERS];                            // pre-stage value of this dimension in every layer
    // ---- forward ----
#pragma unroll 1
    for (int l = L - 1; l >= 0; --l) {
        const GfLayerC<T>& c = g.layers[l];
        if (c.has_offset) xj -= prow[(int64_t)(c.raw_off + j) * sj];
        if (c.hh_iter > 0) {
            ex[(fX + j) * 32] = xj;
            for (int i = 0; i < c.hh_iter; ++i) ex[(fH + i * D + j) * 32] = prow[(int64_t)(c.raw_hh() + i * D + j) * sj];
            fb_bar(bar_id, 32 * D);
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
                const T cc = T(2) * dot / nrm;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) X[jj] = fma(-cc, w[jj], X[jj]);
            }
#pragma unroll
            for (int jj = 0; jj < D; ++jj) xj = (jj == j) ? X[jj] : xj;
            fb_bar(bar_id, 32 * D);            // everybody has read the exchange
        }
        vsave[l] = xj;
        if (live) {
            const MixView<T> mv = regulate_to_slots<T>(c, c.K, j, prow, sj, slots, false);
            T y, logd;
            gf_eval_logpdf<T>(mv, c.inv_type, xj, y, logd);
            xj = y;
            ld_acc += logd;
        }
    }
    // ---- forward outputs: base point, logdet, log N(z) ----
    if (live && a.out != nullptr) a.out[row * a.ld_out + j] = xj;
    if (a.logdet_out != nullptr || a.logbase_out != nullptr) {
        ex[(fX + j) * 32] = ld_acc;
        ex[(fXB + j) * 32] = xj * xj;
        fb_bar(bar_id, 32 * D);
        if (j == 0 && live) {
            T ld = 0, zsq = 0;
#pragma unroll
            for (int jj = 0; jj < D; ++jj) { ld += ex[(fX + jj) * 32]; zsq += ex[(fXB + jj) * 32]; }
            if (!finite_(ld)) status_add(a.status, JF_STATUS_NONFINITE, 1);
            if (a.logdet_out != nullptr) a.logdet_out[row] = ld;
            if (a.logbase_out != nullptr) a.logbase_out[row] = -T(0.5) * zsq - T(D) * T(kLogSqrt2Pi);
        }
        fb_bar(bar_id, 32 * D);
    }
    // ---- backward ----
    const T gr = g.grad_logp ? g.grad_logp[row] : T(1);
    T xb = -xj * gr;                                   // d/dz_j of sum_j log N(z_j)
    int n_bad = 0;
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
        const GfLayerC<T>& c = g.layers[l];
        const T v = vsave[l];
        if (live) {
            const T Gs = bwd_regulate<T>(c, c.K, j, prow, sj, slots);
            T vbar;
            if (!gf_elem_backward<T>(c, c.K, j, v, xb, gr, Gs, slots, grow, sj, vbar)) {
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
            for (int i = 0; i < c.hh_iter; ++i) ex[(fH + i * D + j) * 32] = prow[(int64_t)(c.raw_hh() + i * D + j) * sj];
            fb_bar(bar_id, 32 * D);
            T V[D], XB[D];
#pragma unroll
            for (int jj = 0; jj < D; ++jj) { V[jj] = ex[(fX + jj) * 32]; XB[jj] = ex[(fXB + jj) * 32]; }
            T v_own = v, xb_own = xb;
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
                const T is = T(1) / s;
                const T ain = -aa;
                const T w_own = ex[(fH + i * D + j) * 32];
                const T ca = T(2) * aa * is, cb = T(2) * bb * is;
                const T xin = fma(-ca, w_own, v_own);
                if (live)
                    grow[(int64_t)(c.raw_hh() + i * D + j) * sj] = -T(2) * is * (bb * xin + ain * xb_own) + T(4) * ain * bb * is * is * w_own;
                xb_own = fma(-cb, w_own, xb_own);
                v_own = xin;
#pragma unroll
                for (int jj = 0; jj < D; ++jj) { V[jj] = fma(-ca, w[jj], V[jj]); XB[jj] = fma(-cb, w[jj], XB[jj]); }
            }
            xb = xb_own;
            fb_bar(bar_id, 32 * D);
        }
        if (c.has_offset && live) grow[(int64_t)(c.raw_off + j) * sj] = -xb;
    }
    if (g.grad_x != nullptr && live) g.grad_x[row * g.ld_gx + j] = xb;
    if (n_bad) status_add(a.status, JF_STATUS_OUT_OF_RANGE, n_bad);
}

}  // namespace jf
