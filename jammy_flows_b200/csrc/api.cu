// C-ABI entry points of libjammy_b200.so (see include/jammy_b200.h).  Host-side dispatch only; all math is in the
// kernels.  Built with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
// exp() coefficients as immediates in this unit: measured, the tcgen05 MLP prologue (tanh) is faster that way, while the
// "g" chain kernels (gf_inst_*.cu) gain 2-5 % from constant-bank operands (profiles/ncu_r01_final.md)
#define JF_EXP_CONST 0
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include "subpdf_kernels.cuh"
#include "gf_launch.cuh"
#include "gfx.cuh"
#include "chart.cuh"
#include "rng.cuh"
#include "mlp_kernels.cuh"
#include "mlp_dmma.cuh"
#include "mlp_i8.cuh"
#include "gf_fb_launch.cuh"
#include "jac_sweep.cuh"
#include "rowwise.cuh"
#include "gf_fused_launch.cuh"
#include "mlp_bwd_launch.cuh"
#include <cstdlib>

using namespace jf;

static std::atomic<int64_t> g_launches{0};

#define JF_CUDA_OK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)

static inline int check_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? JF_OK : (int)e;
}

// ---------------------------------------------------------------------------------------------------------------------
// jf_subpdf_apply
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
static void fill_common(SubPdfArgs<T>& a, const JfSubPdfDesc* desc, const void* in, int64_t ld_in, const void* params,
                        int64_t sp, int64_t sr, const void* logdet_in, void* logdet_out, const void* logbase_in,
                        void* logbase_out, void* out, int64_t ld_out, void* emb_out, int64_t ld_emb, int64_t B,
                        int64_t* status) {
    a.n_layers = desc->n_layers;
    a.d = desc->dim;
    a.B = B;
    a.in = (const T*)in; a.ld_in = ld_in;
    a.out = (T*)out; a.ld_out = ld_out;
    a.params = (const T*)params; a.sj = sp; a.sr = sr;
    a.logdet_in = (const T*)logdet_in; a.logdet_out = (T*)logdet_out;
    a.logbase_in = (const T*)logbase_in; a.logbase_out = (T*)logbase_out;
    a.emb_out = (T*)emb_out; a.ld_emb = ld_emb;
    a.status = status;
    a.tab_total = 0;
}

// "g" layer with any non-default option -> general chain kernel (csrc/gfx.cuh)
static bool gf_layer_is_default(const JfLayerDesc& L) {
    return L.rotation_mode == JF_ROT_HOUSEHOLDER && L.width_mode == JF_WIDTH_SMOOTH && !L.width_clamp && !L.skew &&
           !L.center_mean && L.stretch == JF_STRETCH_CLASSIC;
}

// number of rotation parameters of a "g" layer, -1: invalid
static int gf_rotation_params(const JfLayerDesc& L, int d) {
    switch (L.rotation_mode) {
        case JF_ROT_HOUSEHOLDER: return L.hh_iter * d;
        case JF_ROT_NONE: return 0;
        case JF_ROT_ANGLES: return d > 1 ? d * (d - 1) / 2 : 0;
        case JF_ROT_CAYLEY: return d > 1 ? (d == 2 ? 1 : -1) : 0;
        case JF_ROT_TRIANGULAR: return d - 1 + d * (d - 1);
        default: return -1;
    }
}

template <typename T>
static int fill_mvn(GfLayerC<T>& c, const JfLayerDesc& L, int d, int tab) {
    if (L.inv_type < JF_COV_IDENTITY || L.inv_type > JF_COV_FULL) return JF_ERR_BAD_DESC;
    if (!(L.w_min > 0) || !(L.w_max > 0)) return JF_ERR_BAD_DESC;
    const int ncov = L.inv_type == JF_COV_IDENTITY ? 0 : (L.inv_type == JF_COV_DIAGONAL_SYMMETRIC ? 1
                     : (L.inv_type == JF_COV_DIAGONAL ? d : d + d * (d - 1) / 2));
    if ((L.has_offset ? d : 0) + ncov != L.n_params) return JF_ERR_BAD_DESC;
    c.kind = 1; c.K = 1; c.d = d; c.hh_iter = 0; c.inv_type = L.inv_type; c.norm_mode = JF_NORM_NONE;
    c.has_offset = L.has_offset; c.raw_off = L.param_offset; c.tab_off = tab;
    c.w_min = (T)L.w_min; c.inv_w_max = (T)(1.0 / L.w_max); c.n_min = 0; c.n_max = 0;
    return JF_OK;
}

template <typename T>
static int apply_gfx(const JfSubPdfDesc* desc, int direction, const SubPdfArgs<T>& a, cudaStream_t st) {
    const int d = desc->dim;
    GfxChainArgs<T> g;
    memset(&g, 0, sizeof(g));
    g.a = a;
    int kmax = 1;
    for (int l = 0; l < desc->n_layers; ++l) {
        const JfLayerDesc& L = desc->layers[l];
        if (L.dim != d) return JF_ERR_BAD_DESC;
        GfxLayerC<T>& c = g.layers[l];
        if (L.kind == JF_LAYER_MVN) {
            const int rc = fill_mvn<T>(c.base, L, d, 0);
            if (rc != JF_OK) return rc;
            continue;
        }
        if (L.kind != JF_LAYER_GF) return JF_ERR_BAD_DESC;
        if (L.K < 1 || L.K > JF_MAX_KDE || L.hh_iter < 0 || L.hh_iter > 4 * JF_MAX_DIM) return JF_ERR_UNSUPPORTED;
        if (L.inv_type < 0 || L.inv_type > 3 || L.norm_mode < 0 || L.norm_mode > 2) return JF_ERR_BAD_DESC;
        if (L.width_mode < JF_WIDTH_SMOOTH || L.width_mode > JF_WIDTH_SOFTPLUS) return JF_ERR_BAD_DESC;
        if (L.stretch != JF_STRETCH_CLASSIC && L.stretch != JF_STRETCH_RQS) return JF_ERR_BAD_DESC;
        if (!(L.w_min > 0)) return JF_ERR_BAD_DESC;
        if (L.width_mode == JF_WIDTH_SMOOTH && L.stretch == JF_STRETCH_CLASSIC && !(L.w_max > 0)) return JF_ERR_BAD_DESC;
        const int n_rot = gf_rotation_params(L, d);
        if (n_rot < 0) return JF_ERR_BAD_DESC;
        const int kd = L.K * d;
        int off = L.param_offset + (L.has_offset ? d : 0);
        c.base.kind = 0; c.base.K = L.K; c.base.d = d; c.base.hh_iter = L.rotation_mode == JF_ROT_HOUSEHOLDER ? L.hh_iter : 0;
        c.base.inv_type = L.inv_type; c.base.norm_mode = L.norm_mode; c.base.has_offset = L.has_offset;
        c.base.raw_off = L.param_offset; c.base.tab_off = 0;
        c.base.w_min = (T)L.w_min; c.base.inv_w_max = (T)(L.w_max > 0 ? 1.0 / L.w_max : 0.0);
        c.base.n_min = (T)L.n_min; c.base.n_max = (T)L.n_max;
        c.rot_mode = L.rotation_mode; c.width_mode = L.width_mode; c.width_clamp = L.width_clamp; c.skew = L.skew;
        c.center_mean = L.center_mean; c.stretch = L.stretch;
        c.clamp_lo = (T)L.clamp_lo; c.clamp_hi = (T)L.clamp_hi;
        c.off_rot = off; off += n_rot;
        if (L.stretch == JF_STRETCH_RQS) {
            if (L.K * 1e-3 > 1.0) return JF_ERR_BAD_DESC;
            c.off_m = off; off += kd;                 // log_widths [d,K]
            c.off_w = off; off += kd;                 // log_heights [d,K]
            c.off_n = off; off += (L.K + 1) * d;      // log_derivatives [d,K+1]
            c.off_s = off; off += 4 * d;              // boundary_points [d,4]
            c.skew = 0; c.center_mean = 0;
        } else {
            if (L.center_mean && L.K < 2) return JF_ERR_BAD_DESC;
            c.off_m = off; off += (L.K - (L.center_mean ? 1 : 0)) * d;
            c.off_w = off; off += kd;
            c.off_n = off; if (L.norm_mode != JF_NORM_NONE) off += kd;
            c.off_s = off; if (L.skew) off += kd;
        }
        if (off - L.param_offset != L.n_params) return JF_ERR_BAD_DESC;
        kmax = L.K > kmax ? L.K : kmax;
    }
    const int rc = launch_gfx<T>(g, direction, kmax, st);
    if (rc != JF_OK) return rc;
    return check_launch();
}

template <typename T>
static int apply_gf(const JfSubPdfDesc* desc, int direction, GfChainArgs<T>& g, cudaStream_t st) {
    const int d = desc->dim;
    if (d < 1 || d > JF_MAX_DIM) return JF_ERR_UNSUPPORTED;
    for (int l = 0; l < desc->n_layers; ++l)
        if (desc->layers[l].kind == JF_LAYER_GF && !gf_layer_is_default(desc->layers[l]))
            return apply_gfx<T>(desc, direction, g.a, st);
    int kmax = 1;
    int tab = 0;
    for (int l = 0; l < desc->n_layers; ++l) {
        const JfLayerDesc& L = desc->layers[l];
        if (L.dim != d) return JF_ERR_BAD_DESC;
        if (!(L.w_min > 0) || !(L.w_max > 0)) return JF_ERR_BAD_DESC;
        GfLayerC<T>& c = g.layers[l];
        if (L.kind == JF_LAYER_MVN) {           // "t": inv_type carries the covariance type
            const int rc = fill_mvn<T>(c, L, d, tab);
            if (rc != JF_OK) return rc;
            continue;
        }
        if (L.kind != JF_LAYER_GF) return JF_ERR_BAD_DESC;
        if (L.K < 1 || L.K > JF_MAX_KDE || L.hh_iter < 0 || L.hh_iter > 4 * JF_MAX_DIM) return JF_ERR_UNSUPPORTED;
        if (L.inv_type < 0 || L.inv_type > 3 || L.norm_mode < 0 || L.norm_mode > 2) return JF_ERR_BAD_DESC;
        const int expect = (L.has_offset ? d : 0) + L.hh_iter * d + (L.norm_mode != JF_NORM_NONE ? 3 : 2) * L.K * d;
        if (expect != L.n_params) return JF_ERR_BAD_DESC;
        c.kind = 0;
        c.K = L.K; c.d = d; c.hh_iter = L.hh_iter; c.inv_type = L.inv_type; c.norm_mode = L.norm_mode;
        c.has_offset = L.has_offset; c.raw_off = L.param_offset; c.tab_off = tab;
        c.w_min = (T)L.w_min; c.inv_w_max = (T)(1.0 / L.w_max); c.n_min = (T)L.n_min; c.n_max = (T)L.n_max;
        tab += c.tab_size();
        kmax = L.K > kmax ? L.K : kmax;
    }
    g.a.tab_total = tab;
    // per-row parameters, log_pdf direction: the (row, dimension)-worker kernel of csrc/gf_fb.cuh in its forward-only mode
    // (coalesced param-major loads, register-resident unrolled mixture).  JF_ROWDIM_FWD=0 keeps the thread-per-row kernel.
    static const bool rowdim = [] { const char* e = getenv("JF_ROWDIM_FWD"); return e == nullptr || atoi(e) != 0; }();
    if (rowdim && direction == JF_DIR_LOGPDF && g.a.sr != 0) {
        GfFbArgs<T> f;
        memset(&f, 0, sizeof(f));
        f.a = g.a;
        int hh_max = 0;
        for (int l = 0; l < desc->n_layers; ++l) {
            f.layers[l] = g.layers[l];
            hh_max = g.layers[l].hh_iter > hh_max ? g.layers[l].hh_iter : hh_max;
        }
        f.kmax = kmax;
        f.hh_max = hh_max;
        const int rc = launch_gf_fwd<T>(f, st);
        if (rc == JF_OK) return check_launch();
        if (rc != JF_ERR_UNSUPPORTED) return rc;          // (too much shared memory for this shape: the thread-per-row kernel)
    }
    const size_t smem = (g.a.sr == 0) ? (size_t)tab * sizeof(T) : 0;
    if (smem > 200 * 1024) return JF_ERR_UNSUPPORTED;
    const int rc = (direction == JF_DIR_LOGPDF) ? launch_gf_dir<T, JF_DIR_LOGPDF>(g, d, kmax, smem, st)
                                                : launch_gf_dir<T, JF_DIR_SAMPLE>(g, d, kmax, smem, st);
    if (rc != JF_OK) return rc;
    return check_launch();
}

// JfSplineDesc -> device constants.  `rel_off`: extra offset added to the descriptor's param_offset.
template <typename T>
static int fill_spline(SplineC<T>& c, const JfSplineDesc& d, int rel_off) {
    if (d.n_bins < 1 || d.n_bins > JF_MAX_BINS) return JF_ERR_UNSUPPORTED;
    if (d.kind < JF_SPLINE_PLAIN || d.kind > JF_SPLINE_CIRCULAR || d.bd_mode < JF_BD_PARAMS || d.bd_mode > JF_BD_PERIODIC)
        return JF_ERR_BAD_DESC;
    if (d.kind == JF_SPLINE_SMOOTH && d.n_bins > 3) return JF_ERR_UNSUPPORTED;     // reference: 2 or 3 bins only
    if (d.kind == JF_SPLINE_CIRCULAR && d.n_bins != 2) return JF_ERR_UNSUPPORTED;
    if (!(d.hi > d.lo) || d.min_w * d.n_bins > 1.0 || d.min_h * d.n_bins > 1.0) return JF_ERR_BAD_DESC;
    // parameter counts implied by the options (rational_quadratic_spline.py:98-157, splines_1d.py:38-94)
    const bool mirror = d.kind == JF_SPLINE_SMOOTH && d.n_bins == 3;
    const int L = d.n_bins - (mirror ? 1 : 0);
    const int zw = d.fix_first ? (d.fix_second ? 2 : 1) : 0, zh = d.fix_first ? 1 : 0;
    int nd;
    if (d.kind == JF_SPLINE_PLAIN) nd = d.bd_mode == JF_BD_FIXED ? d.n_bins - 1 : (d.bd_mode == JF_BD_PERIODIC ? d.n_bins : d.n_bins + 1);
    else if (d.kind == JF_SPLINE_SMOOTH) nd = d.bd_mode == JF_BD_FIXED ? 0 : 2;
    else nd = 0;
    if (d.n_w != L - zw || d.n_h != L - zh || d.n_d != nd || d.n_w < 0) return JF_ERR_BAD_DESC;
    c.kind = d.kind; c.n_bins = d.n_bins; c.n_w = d.n_w; c.n_h = d.n_h; c.n_d = d.n_d;
    c.fix_first = d.fix_first; c.fix_second = d.fix_second; c.indep = d.indep; c.bd_mode = d.bd_mode;
    c.natural_direction = d.natural_direction; c.raw_off = d.param_offset + rel_off; c.pad_ = 0;
    c.lo = (T)d.lo; c.hi = (T)d.hi; c.min_w = (T)d.min_w; c.min_h = (T)d.min_h; c.min_d = (T)d.min_d;
    c.bd_fixed = (T)d.bd_fixed;
    c.ln_max_ratio = T(-1);
    if (d.max_ratio > 0.0) {
        if (d.n_bins < 2) return JF_ERR_BAD_DESC;
        const double l = (log(d.max_ratio) - log((double)(d.n_bins - 1))) / 2.0;
        if (!(l > 0.0)) return JF_ERR_BAD_DESC;
        c.ln_max_ratio = (T)l;
    }
    return JF_OK;
}

template <typename T>
static int fill_s2(const JfSubPdfDesc* desc, S2Args<T>& g) {
    if (desc->dim != 2) return JF_ERR_UNSUPPORTED;
    int n_sp = 0;
    for (int l = 0; l < desc->n_layers; ++l) {
        const JfLayerDesc& L = desc->layers[l];
        if (L.kind != JF_LAYER_FVM && L.kind != JF_LAYER_EXPMAP) return JF_ERR_UNSUPPORTED;
        if (L.first != 0 && l != 0) return JF_ERR_BAD_DESC;   // only layer 0 may carry the chart
        FvmLayerC& c = g.layers[l];
        memset(&c, 0, sizeof(c));
        c.kind = L.kind;
        c.add_rotation = L.hh_iter > 0; c.hh_iter = L.hh_iter; c.first = L.first; c.raw_off = L.param_offset;
        int n_hh = L.hh_iter > 0 ? L.hh_iter * 3 : 0;
        c.rot_mode = JF_ROT_HOUSEHOLDER; c.n_rot = n_hh; c.kappa_mode = JF_KAPPA_DIRECT_LOG; c.kappa_clamp = 0;
        if (L.kind == JF_LAYER_FVM) {
            // rotation modes of sphere_base.py:79-91 and the kappa link functions of fvm_2d.py:108-138
            if (L.rotation_mode == JF_ROT_ANGLES || L.rotation_mode == JF_ROT_XYZ || L.rotation_mode == JF_ROT_QUATERNION) {
                c.add_rotation = 1; c.hh_iter = 0; c.rot_mode = L.rotation_mode;
                n_hh = L.rotation_mode == JF_ROT_QUATERNION ? 4 : 3;
                c.n_rot = n_hh;
            } else if (L.rotation_mode != JF_ROT_HOUSEHOLDER) {
                return JF_ERR_UNSUPPORTED;
            }
            if (L.width_mode < JF_KAPPA_DIRECT_LOG || L.width_mode > JF_KAPPA_QUATVEC_SQUARED) return JF_ERR_BAD_DESC;
            c.kappa_mode = L.width_mode; c.kappa_clamp = L.width_clamp != 0;
            c.extra_rot = L.skew != 0; c.identity_region = L.clamp_lo;
            if (!(L.clamp_lo >= 0.0 && L.clamp_lo < 1.0)) return JF_ERR_BAD_DESC;
            if ((c.kappa_mode == JF_KAPPA_MU || c.kappa_mode == JF_KAPPA_MU_SQUARED) && c.rot_mode != JF_ROT_XYZ) return JF_ERR_BAD_DESC;
            if ((c.kappa_mode == JF_KAPPA_QUATVEC || c.kappa_mode == JF_KAPPA_QUATVEC_SQUARED) && c.rot_mode != JF_ROT_QUATERNION)
                return JF_ERR_BAD_DESC;
            if (c.kappa_mode >= JF_KAPPA_MU) n_hh -= 1;      // no kappa parameter of its own: "expect = n_hh + 1" below
            c.z_sign = L.z_sign; c.min_kappa = L.min_kappa;
            if (L.n_vertical < 0 || L.n_circular < 0 || L.n_vertical + L.n_circular > JF_MAX_NESTED) return JF_ERR_BAD_DESC;
            if (n_sp + L.n_vertical + L.n_circular > kS2MaxSplines) return JF_ERR_UNSUPPORTED;
            c.v_first = n_sp; c.n_vertical = L.n_vertical;
            c.c_first = n_sp + L.n_vertical; c.n_circular = L.n_circular;
            int expect = n_hh + 1, off = 0;
            for (int i = 0; i < L.n_vertical + L.n_circular; ++i) {
                if (L.spline[i].param_offset != off) return JF_ERR_BAD_DESC;
                const int rc = fill_spline<T>(g.splines[n_sp], L.spline[i], 0);
                if (rc != JF_OK) return rc;
                off += g.splines[n_sp].n_params();
                ++n_sp;
            }
            expect += off;
            if (expect != L.n_params) return JF_ERR_BAD_DESC;
        } else {
            if (L.K < 1 || L.K > kMaxExpComp) return JF_ERR_UNSUPPORTED;
            if (L.inv_type < JF_POT_EXPONENTIAL || L.inv_type > JF_POT_SPLINES) return JF_ERR_BAD_DESC;
            const int n_pot = L.inv_type == JF_POT_EXPONENTIAL ? 5 : (L.inv_type == JF_POT_SPLINES ? 4 + 3 * kVSplineBins + 1 : 4);
            if (n_hh + n_pot * L.K != L.n_params) return JF_ERR_BAD_DESC;
            c.pot = L.inv_type;
            c.K = L.K; c.natural_direction = L.natural_direction; c.max_iter = L.max_iter > 0 ? L.max_iter : 1000;
        }
    }
    return JF_OK;
}

template <typename T>
static int apply_s2(const JfSubPdfDesc* desc, int direction, S2Args<T>& g, cudaStream_t st) {
    const int rc = fill_s2<T>(desc, g);
    if (rc != JF_OK) return rc;
    const int threads = 128;
    const int64_t blocks = (g.a.B + threads - 1) / threads;
    if (direction == JF_DIR_LOGPDF) s2_chain_kernel<T, JF_DIR_LOGPDF><<<(unsigned)blocks, threads, 0, st>>>(g);
    else s2_chain_kernel<T, JF_DIR_SAMPLE><<<(unsigned)blocks, threads, 0, st>>>(g);
    return check_launch();
}

// one-dimensional sub-pdfs: interval ("r") and circle ("o", "m")
template <typename T>
static int fill_chain1(const JfSubPdfDesc* desc, Chain1Args<T>& g) {
    if (desc->dim != 1) return JF_ERR_UNSUPPORTED;
    g.manifold = desc->manifold;
    for (int l = 0; l < desc->n_layers; ++l) {
        const JfLayerDesc& L = desc->layers[l];
        Layer1C<T>& c = g.layers[l];
        memset(&c, 0, sizeof(c));
        if (L.first != 0 && l != 0) return JF_ERR_BAD_DESC;
        c.kind = L.kind; c.first = L.first; c.raw_off = L.param_offset; c.natural_direction = L.natural_direction;
        if (desc->manifold == 'i') {
            if (L.kind != JF_LAYER_RQS) return JF_ERR_UNSUPPORTED;
            if (!(L.hi > L.lo)) return JF_ERR_BAD_DESC;
            c.lo = (T)L.lo; c.hi = (T)L.hi;
            const int rc = fill_spline<T>(c.sp, L.spline[0], 0);
            if (rc != JF_OK) return rc;
            if (c.sp.n_params() != L.n_params) return JF_ERR_BAD_DESC;
        } else {
            if (L.hh_iter < 0 || L.hh_iter > 8) return JF_ERR_UNSUPPORTED;
            c.hh_iter = L.hh_iter;
            if (L.kind == JF_LAYER_S1SPLINE) {
                const int rc = fill_spline<T>(c.sp, L.spline[0], 0);
                if (rc != JF_OK) return rc;
                if (c.sp.n_params() + 2 * L.hh_iter != L.n_params) return JF_ERR_BAD_DESC;
            } else if (L.kind == JF_LAYER_MOEBIUS) {
                if (L.K < 1 || L.K > kMaxMoebius) return JF_ERR_UNSUPPORTED;
                if (4 * L.K + 2 * L.hh_iter != L.n_params) return JF_ERR_BAD_DESC;
                c.K = L.K;
            } else {
                return JF_ERR_UNSUPPORTED;
            }
        }
    }
    return JF_OK;
}

template <typename T>
static int apply_chain1(const JfSubPdfDesc* desc, int direction, Chain1Args<T>& g, cudaStream_t st) {
    const int rc = fill_chain1<T>(desc, g);
    if (rc != JF_OK) return rc;
    const int threads = 128;
    const int64_t blocks = (g.a.B + threads - 1) / threads;
    if (direction == JF_DIR_LOGPDF) chain1_kernel<T, JF_DIR_LOGPDF><<<(unsigned)blocks, threads, 0, st>>>(g);
    else chain1_kernel<T, JF_DIR_SAMPLE><<<(unsigned)blocks, threads, 0, st>>>(g);
    return check_launch();
}

template <typename T>
static int subpdf_apply_t(const JfSubPdfDesc* desc, int direction, const void* in, int64_t ld_in, const void* params,
                          int64_t sp, int64_t sr, const void* logdet_in, void* logdet_out, const void* logbase_in,
                          void* logbase_out, void* out, int64_t ld_out, void* emb_out, int64_t ld_emb, int64_t B,
                          int64_t* status, cudaStream_t st) {
    if (desc->manifold == 'e') {
        GfChainArgs<T> g;
        fill_common<T>(g.a, desc, in, ld_in, params, sp, sr, logdet_in, logdet_out, logbase_in, logbase_out, out, ld_out,
                       emb_out, ld_emb, B, status);
        return apply_gf<T>(desc, direction, g, st);
    }
    if ((desc->manifold == 's' && desc->dim == 1) || desc->manifold == 'i') {
        Chain1Args<T> g;
        fill_common<T>(g.a, desc, in, ld_in, params, sp, sr, logdet_in, logdet_out, logbase_in, logbase_out, out, ld_out,
                       emb_out, ld_emb, B, status);
        return apply_chain1<T>(desc, direction, g, st);
    }
    if (desc->manifold == 's') {
        S2Args<T> g;
        fill_common<T>(g.a, desc, in, ld_in, params, sp, sr, logdet_in, logdet_out, logbase_in, logbase_out, out, ld_out,
                       emb_out, ld_emb, B, status);
        return apply_s2<T>(desc, direction, g, st);
    }
    return JF_ERR_UNSUPPORTED;
}

extern "C" int jf_subpdf_apply(const JfSubPdfDesc* desc, int dtype, int direction, const void* in, int64_t ld_in,
                               const void* params, int64_t p_stride_param, int64_t p_stride_row, const void* logdet_in,
                               void* logdet_out, const void* logbase_in, void* logbase_out, void* out, int64_t ld_out,
                               void* emb_out, int64_t ld_emb, int64_t B, int64_t* status, void* stream) {
    if (desc == nullptr || in == nullptr || out == nullptr) return JF_ERR_BAD_ARG;
    if (desc->n_layers < 1 || desc->n_layers > JF_MAX_LAYERS) return JF_ERR_BAD_DESC;
    if (direction != JF_DIR_LOGPDF && direction != JF_DIR_SAMPLE) return JF_ERR_BAD_ARG;
    if (desc->n_params > 0 && params == nullptr) return JF_ERR_BAD_ARG;
    if (B < 0 || B > (int64_t)2147483647 * 256) return JF_ERR_BAD_ARG;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == JF_F64)
        return subpdf_apply_t<double>(desc, direction, in, ld_in, params, p_stride_param, p_stride_row, logdet_in,
                                      logdet_out, logbase_in, logbase_out, out, ld_out, emb_out, ld_emb, B, status, st);
    if (dtype == JF_F32)
        return subpdf_apply_t<float>(desc, direction, in, ld_in, params, p_stride_param, p_stride_row, logdet_in,
                                     logdet_out, logbase_in, logbase_out, out, ld_out, emb_out, ld_emb, B, status, st);
    return JF_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------------------------------------
// jf_subpdf_forward_backward / jf_subpdf_backward: log_pdf of a Euclidean "g" sub-pdf and its per-row parameter gradients
// in one kernel (csrc/gf_fb.cuh)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
static int subpdf_fb_t(const JfSubPdfDesc* desc, const void* x, int64_t ld_x, const void* params, int64_t sp,
                       int64_t sr, const void* grad_logp, void* grad_params, void* grad_x, int64_t ld_gx, void* base_out,
                       int64_t ld_out, void* logdet_out, void* logbase_out, int64_t B, int64_t* status, cudaStream_t st,
                       const void* grad_out_x = nullptr, int64_t ld_go = 0, int mode = 0) {
    GfFbArgs<T> g;
    fill_common<T>(g.a, desc, x, ld_x, params, sp, sr, nullptr, logdet_out, nullptr, logbase_out, base_out, ld_out, nullptr, 0,
                   B, status);
    g.grad_logp = (const T*)grad_logp;
    g.grad_params = (T*)grad_params;
    g.grad_x = (T*)grad_x;
    g.ld_gx = ld_gx;
    g.grad_out_x = (const T*)grad_out_x;
    g.ld_go = ld_go;
    const int d = desc->dim;
    if (d < 1 || d > JF_MAX_DIM) return JF_ERR_UNSUPPORTED;
    int kmax = 1, hh_max = 0;
    for (int l = 0; l < desc->n_layers; ++l) {
        const JfLayerDesc& L = desc->layers[l];
        if (L.kind == JF_LAYER_MVN && L.dim == d) {               // "t": affine layer, reverse pass in fb_mvn_backward
            const int rc = fill_mvn<T>(g.layers[l], L, d, 0);
            if (rc != JF_OK) return rc;
            continue;
        }
        if (L.kind != JF_LAYER_GF || L.dim != d) return JF_ERR_UNSUPPORTED;
        // the closed-form backward covers the default options and rotation_mode "none" (no reflections at all)
        JfLayerDesc Lr = L;
        if (Lr.rotation_mode == JF_ROT_NONE && Lr.hh_iter == 0) Lr.rotation_mode = JF_ROT_HOUSEHOLDER;
        if (!gf_layer_is_default(Lr)) return JF_ERR_UNSUPPORTED;
        if (L.K < 1 || L.K > JF_MAX_KDE || L.hh_iter < 0 || L.hh_iter > 4 * JF_MAX_DIM) return JF_ERR_UNSUPPORTED;
        if (L.inv_type == JF_INV_FULL_PADE || L.inv_type == JF_INV_PARTLY_CRUDE) return JF_ERR_UNSUPPORTED;
        const int expect = (L.has_offset ? d : 0) + L.hh_iter * d + (L.norm_mode != JF_NORM_NONE ? 3 : 2) * L.K * d;
        if (expect != L.n_params) return JF_ERR_BAD_DESC;
        GfLayerC<T>& c = g.layers[l];
        c.K = L.K; c.d = d; c.hh_iter = L.hh_iter; c.inv_type = L.inv_type; c.norm_mode = L.norm_mode;
        c.kind = 0;
        c.has_offset = L.has_offset; c.raw_off = L.param_offset; c.tab_off = 0;
        c.w_min = (T)L.w_min; c.inv_w_max = (T)(1.0 / L.w_max); c.n_min = (T)L.n_min; c.n_max = (T)L.n_max;
        kmax = L.K > kmax ? L.K : kmax;
        hh_max = L.hh_iter > hh_max ? L.hh_iter : hh_max;
    }
    g.kmax = kmax;
    g.hh_max = hh_max;
    const int rc = mode == 0 ? launch_gf_fb<T>(g, st) : launch_gf_sbwd<T>(g, st);
    if (rc != JF_OK) return rc;
    return check_launch();
}

static int subpdf_fb_checks(const JfSubPdfDesc* desc, const void* x, const void* params, const void* grad_params,
                            int64_t p_stride_row, int64_t B) {
    if (desc == nullptr || x == nullptr || params == nullptr || grad_params == nullptr) return JF_ERR_BAD_ARG;
    if (desc->n_layers < 1 || desc->n_layers > JF_MAX_LAYERS) return JF_ERR_BAD_DESC;
    if (desc->manifold != 'e') return JF_ERR_UNSUPPORTED;
    if (p_stride_row == 0) return JF_ERR_UNSUPPORTED;      // shared (permanent) parameters: expand them to per-row form
    if (B < 0) return JF_ERR_BAD_ARG;
    return JF_OK;
}

extern "C" int jf_subpdf_backward(const JfSubPdfDesc* desc, int dtype, const void* x, int64_t ld_x, const void* params,
                                  int64_t p_stride_param, int64_t p_stride_row, const void* grad_logp, void* grad_params,
                                  int64_t B, int64_t* status, void* stream) {
    return jf_subpdf_forward_backward(desc, dtype, x, ld_x, params, p_stride_param, p_stride_row, grad_logp, grad_params,
                                      nullptr, 0, nullptr, 0, nullptr, nullptr, B, status, stream);
}

extern "C" int jf_subpdf_forward_backward(const JfSubPdfDesc* desc, int dtype, const void* x, int64_t ld_x,
                                          const void* params, int64_t p_stride_param, int64_t p_stride_row,
                                          const void* grad_logp, void* grad_params, void* grad_x, int64_t ld_gx,
                                          void* base_out, int64_t ld_out, void* logdet_out, void* logbase_out, int64_t B,
                                          int64_t* status, void* stream) {
    const int rc = subpdf_fb_checks(desc, x, params, grad_params, p_stride_row, B);
    if (rc != JF_OK) return rc;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == JF_F64)
        return subpdf_fb_t<double>(desc, x, ld_x, params, p_stride_param, p_stride_row, grad_logp, grad_params, grad_x, ld_gx,
                                   base_out, ld_out, logdet_out, logbase_out, B, status, st);
    if (dtype == JF_F32)
        return subpdf_fb_t<float>(desc, x, ld_x, params, p_stride_param, p_stride_row, grad_logp, grad_params, grad_x, ld_gx,
                                  base_out, ld_out, logdet_out, logbase_out, B, status, st);
    return JF_ERR_BAD_ARG;
}

extern "C" int jf_subpdf_sample_backward(const JfSubPdfDesc* desc, int dtype, const void* x, int64_t ld_x,
                                         const void* params, int64_t p_stride_param, int64_t p_stride_row,
                                         const void* grad_x, int64_t ld_gx, const void* grad_logp, void* grad_params,
                                         void* grad_z, int64_t ld_gz, int64_t B, int64_t* status, void* stream) {
    const int rc = subpdf_fb_checks(desc, x, params, grad_params, p_stride_row, B);
    if (rc != JF_OK) return rc;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == JF_F64)
        return subpdf_fb_t<double>(desc, x, ld_x, params, p_stride_param, p_stride_row, grad_logp, grad_params, grad_z, ld_gz,
                                   nullptr, 0, nullptr, nullptr, B, status, st, grad_x, ld_gx, 1);
    if (dtype == JF_F32)
        return subpdf_fb_t<float>(desc, x, ld_x, params, p_stride_param, p_stride_row, grad_logp, grad_params, grad_z, ld_gz,
                                  nullptr, 0, nullptr, nullptr, B, status, st, grad_x, ld_gx, 1);
    return JF_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------------------------------------
// jf_subpdf_jacobian: per-row Jacobian of the log_pdf of a non-Euclidean sub-pdf (forward-mode sweep, csrc/jac_sweep.cuh)
// ---------------------------------------------------------------------------------------------------------------------
template <typename F>
static int subpdf_jacobian_t(const JfSubPdfDesc* desc, const void* x, int64_t ld_x, const void* params, int64_t sp, int64_t sr,
                             void* jac, int64_t jac_sj, void* jx, int64_t ld_jx, void* jbase, int64_t B, cudaStream_t st) {
    JacIO<F> io;
    io.n_params = desc->n_params; io.d = desc->dim; io.B = B;
    io.in = (const F*)x; io.ld_in = ld_x;
    io.params = (const F*)params; io.sj = sp; io.sr = sr;
    io.jac = (F*)jac; io.jac_sj = jac_sj;
    io.jx = (F*)jx; io.ld_jx = ld_jx;
    io.jbase = (F*)jbase;
    const int threads = 128;
    const dim3 grid((unsigned)((B + threads - 1) / threads), (unsigned)(desc->n_params + (jx != nullptr ? desc->dim : 0)));
    if (grid.y == 0) return JF_OK;
    if (desc->manifold == 's' && desc->dim == 2) {
        S2Args<Dual> g;
        memset(&g.a, 0, sizeof(g.a));
        g.a.n_layers = desc->n_layers; g.a.d = desc->dim; g.a.B = B;
        const int rc = fill_s2<Dual>(desc, g);
        if (rc != JF_OK) return rc;
        for (int l = 0; l < desc->n_layers; ++l)      // the inverse of "v" in its non-natural direction is an iteration
            if (g.layers[l].kind == JF_LAYER_EXPMAP && g.layers[l].natural_direction != 0) return JF_ERR_UNSUPPORTED;
        s2_jac_kernel<F><<<grid, threads, 0, st>>>(io, g);
        return check_launch();
    }
    if ((desc->manifold == 's' && desc->dim == 1) || desc->manifold == 'i') {
        Chain1Args<Dual> g;
        memset(&g.a, 0, sizeof(g.a));
        g.a.n_layers = desc->n_layers; g.a.d = desc->dim; g.a.B = B;
        const int rc = fill_chain1<Dual>(desc, g);
        if (rc != JF_OK) return rc;
        chain1_jac_kernel<F><<<grid, threads, 0, st>>>(io, g);
        return check_launch();
    }
    return JF_ERR_UNSUPPORTED;
}

extern "C" int jf_subpdf_jacobian(const JfSubPdfDesc* desc, int dtype, const void* x, int64_t ld_x, const void* params,
                                  int64_t p_stride_param, int64_t p_stride_row, void* jac_params, int64_t jac_stride_param,
                                  void* jac_x, int64_t ld_jx, void* jac_base, int64_t B, int64_t* status, void* stream) {
    (void)status;
    if (jac_base != nullptr && jac_x == nullptr) return JF_ERR_BAD_ARG;
    if (desc == nullptr || x == nullptr) return JF_ERR_BAD_ARG;
    if (desc->n_layers < 1 || desc->n_layers > JF_MAX_LAYERS) return JF_ERR_BAD_DESC;
    if (desc->manifold == 'e') return JF_ERR_UNSUPPORTED;      // Euclidean chains: jf_subpdf_forward_backward
    if (desc->n_params < 0 || desc->n_params > kJacMaxParams) return JF_ERR_UNSUPPORTED;
    if (desc->n_params > 0 && (params == nullptr || jac_params == nullptr)) return JF_ERR_BAD_ARG;
    if (B < 0) return JF_ERR_BAD_ARG;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == JF_F64)
        return subpdf_jacobian_t<double>(desc, x, ld_x, params, p_stride_param, p_stride_row, jac_params, jac_stride_param,
                                         jac_x, ld_jx, jac_base, B, st);
    if (dtype == JF_F32)
        return subpdf_jacobian_t<float>(desc, x, ld_x, params, p_stride_param, p_stride_row, jac_params, jac_stride_param,
                                        jac_x, ld_jx, jac_base, B, st);
    return JF_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------------------------------------
// jf_mlp_forward
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int TM>
static int launch_mlp(const MlpArgs<T>& m, size_t smem, cudaStream_t st) {
    JF_CUDA_OK(cudaFuncSetAttribute(mlp_kernel<T, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (m.B + TM - 1) / TM;
    mlp_kernel<T, TM><<<(unsigned)blocks, 256, smem, st>>>(m);
    return check_launch();
}

// fp64, one hidden layer of <= 128 units: the DMMA kernel.  Returns JF_ERR_UNSUPPORTED when the shape does not fit.
template <int HP>
static int launch_mlp2_dmma(const MlpArgs<double>& m, cudaStream_t st) {
    const size_t smem = mlp2_dmma_smem(m.dims[0], HP);
    if (smem > 220 * 1024) return JF_ERR_UNSUPPORTED;
    JF_CUDA_OK(cudaFuncSetAttribute(mlp2_dmma_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (m.B + kDmmaRows - 1) / kDmmaRows;
    mlp2_dmma_kernel<HP><<<(unsigned)blocks, kDmmaThreads, smem, st>>>(m);
    return check_launch();
}
static int try_mlp2_dmma(const MlpArgs<double>& m, cudaStream_t st) {
    if (m.n_linear != 2) return JF_ERR_UNSUPPORTED;
    const int Kin = m.dims[0], H = m.dims[1];
    if (H > 128 || (H & 1) || Kin > 256) return JF_ERR_UNSUPPORTED;
    // the W2 / b2 tiles are streamed with 16-byte cp.async: sub-views of a flat parameter vector may start at an odd
    // element, those take the generic kernel
    if ((((uintptr_t)m.wt[1]) | ((uintptr_t)m.bias[1])) & 15) return JF_ERR_UNSUPPORTED;
    if (H <= 32) return launch_mlp2_dmma<32>(m, st);
    if (H <= 64) return launch_mlp2_dmma<64>(m, st);
    return launch_mlp2_dmma<128>(m, st);
}
template <typename T>
static int try_mlp2_dmma(const MlpArgs<T>&, cudaStream_t) { return JF_ERR_UNSUPPORTED; }

// fp64, hidden width 128, <= 16 inputs: the tcgen05 (int8-sliced, exact) kernel.  Needs the caller's workspace for the
// pre-sliced last-layer weights; `prepared` skips the slicing pass (same weights as the previous call on this stream).
#ifndef JF_I8_NS
#define JF_I8_NS 7
#endif
constexpr int kI8NS = JF_I8_NS;      // fp64: 7 int8 slices (54 fractional bits); 6 (46 bits) is an experiment variant
constexpr int kI8NSF32 = 4;   // fp32: 4 int8 slices (30 fractional bits > the 24-bit significand)
static int i8_tn() {       // output-tile width: 64 (default; N = 32 MMAs run at the same 32 cycles: A-read bound) or 32
    static const int tn = [] { const char* e = getenv("JF_I8_TN"); return (e != nullptr && atoi(e) == 32) ? 32 : 64; }();
    return tn;
}
static bool i8_eligible(const JfMlpDesc* d, int dtype) {
    static const bool off = [] { const char* e = getenv("JF_MLP_PATH"); return e != nullptr && strcmp(e, "dmma") == 0; }();
    if (off || d->n_linear != 2 || d->dims[1] != kI8H || d->dims[0] < 1 || d->dims[2] < 1) return false;
    if (dtype == JF_F64) return d->dims[0] <= kI8MaxKin;
    if (dtype == JF_F32) return d->dims[0] <= kI8MaxKinF32;
    return false;
}
static int64_t i8_ws_bytes(int N, int dtype) {
    if (dtype == JF_F64) {
        const int64_t a = i8_prep_bytes<kI8NS, 32>(N), b = i8_prep_bytes<kI8NS, 64>(N);
        return a > b ? a : b;
    }
    const int64_t a = i8_prep_bytes<kI8NSF32, 32>(N), b = i8_prep_bytes<kI8NSF32, 64>(N);
    return a > b ? a : b;
}
template <typename T, int NS, int TN, int KR>
static int launch_mlp2_i8_tn(const MlpArgs<T>& m, void* ws, int prepared, cudaStream_t st) {
    using Cfg = I8Cfg<NS, TN>;
    const int N = m.dims[2], n_tiles = (N + TN - 1) / TN;
    if (!prepared) {
        mlp_i8_prep_kernel<T, NS, TN><<<n_tiles, 128, 0, st>>>(m.wt[1], N, (unsigned char*)ws);
        const int rc = check_launch();
        if (rc != JF_OK) return rc;
    }
    int dev = 0, sms = 0, smem_max = 0;
    JF_CUDA_OK(cudaGetDevice(&dev));
    JF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    JF_CUDA_OK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int n_slots = kI8MaxSlots;
    while (n_slots > NS && Cfg::smem_bytes(m.dims[0], n_slots, (int)sizeof(T)) > smem_max) --n_slots;
    if (n_slots < NS + 1) return JF_ERR_UNSUPPORTED;
    const int smem = Cfg::smem_bytes(m.dims[0], n_slots, (int)sizeof(T));
    JF_CUDA_OK(cudaFuncSetAttribute(mlp2_i8_kernel<T, NS, TN, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t blocks = (m.B + kI8Rows - 1) / kI8Rows;
    const unsigned grid = (unsigned)(blocks < sms ? blocks : sms);     // persistent: one CTA per SM
    mlp2_i8_kernel<T, NS, TN, KR><<<grid, kI8Threads, smem, st>>>(m, (const unsigned char*)ws, n_slots);
    return check_launch();
}
static int launch_mlp2_i8(const MlpArgs<double>& m, void* ws, int prepared, cudaStream_t st) {
    if (i8_tn() == 32) return launch_mlp2_i8_tn<double, kI8NS, 32, 16>(m, ws, prepared, st);
    // <= 8 inputs (cfg2: 4 and 7): half the input registers, 17 % less prologue time (register pressure at 128 regs/thread)
    return m.dims[0] <= 8 ? launch_mlp2_i8_tn<double, kI8NS, 64, 8>(m, ws, prepared, st)
                          : launch_mlp2_i8_tn<double, kI8NS, 64, 16>(m, ws, prepared, st);
}
static int launch_mlp2_i8(const MlpArgs<float>& m, void* ws, int prepared, cudaStream_t st) {
    return i8_tn() == 64 ? launch_mlp2_i8_tn<float, kI8NSF32, 64, 1>(m, ws, prepared, st)
                         : launch_mlp2_i8_tn<float, kI8NSF32, 32, 1>(m, ws, prepared, st);
}

template <typename T>
static int mlp_forward_t(const JfMlpDesc* desc, const void* const* seg_ptrs, const int64_t* seg_ld,
                         const void* const* weights, const void* const* biases, void* out, int64_t so_p, int64_t so_r,
                         int64_t B, cudaStream_t st, void* ws = nullptr, int64_t ws_bytes = 0, int prepared = 0,
                         int accumulate = 0) {
    MlpArgs<T> m;
    memset(&m, 0, sizeof(m));
    m.acc = accumulate ? 1 : 0;
    m.n_linear = desc->n_linear;
    int maxd = 0, in_sum = 0;
    for (int l = 0; l <= desc->n_linear; ++l) {
        if (desc->dims[l] < 1) return JF_ERR_BAD_DESC;
        m.dims[l] = desc->dims[l];
        if (l < desc->n_linear) maxd = desc->dims[l] > maxd ? desc->dims[l] : maxd;   // tiles hold input + hidden acts
    }
    m.n_segments = desc->n_segments;
    for (int s = 0; s < desc->n_segments; ++s) {
        m.seg_cols[s] = desc->seg_cols[s];
        m.seg_ptr[s] = (const T*)seg_ptrs[s];
        m.seg_ld[s] = seg_ld[s];
        in_sum += desc->seg_cols[s];
        if (seg_ptrs[s] == nullptr || desc->seg_cols[s] < 1) return JF_ERR_BAD_ARG;
    }
    if (in_sum != desc->dims[0]) return JF_ERR_BAD_DESC;
    for (int l = 0; l < desc->n_linear; ++l) {
        m.wt[l] = (const T*)weights[l];
        m.bias[l] = (const T*)biases[l];
        if (m.wt[l] == nullptr || m.bias[l] == nullptr) return JF_ERR_BAD_ARG;
    }
    m.out = (T*)out; m.so_p = so_p; m.so_r = so_r; m.B = B;
    m.lda = maxd | 1;
    if (m.n_linear == 1 && m.dims[0] <= kExpandK && m.dims[1] >= 256 && so_p == 1) {
        // narrow input, wide row-major output (the U half of a factorised layer): write-bound expand kernel
        int dev = 0, sms = 148;
        JF_CUDA_OK(cudaGetDevice(&dev));
        JF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int64_t col_tiles = (m.dims[1] + 511) / 512, row_tiles = (B + kExpandRows - 1) / kExpandRows;
        int64_t gy = ((int64_t)sms * 6 + col_tiles - 1) / col_tiles;      // ~6 CTAs per SM in total, each walks row tiles
        if (gy > row_tiles) gy = row_tiles;
        if (gy > 65535) gy = 65535;
        dim3 grid((unsigned)col_tiles, (unsigned)gy);
        if (m.dims[0] <= 4) mlp_expand_kernel<T, 4><<<grid, 256, 0, st>>>(m);
        else if (m.dims[0] <= 8) mlp_expand_kernel<T, 8><<<grid, 256, 0, st>>>(m);
        else mlp_expand_kernel<T, 16><<<grid, 256, 0, st>>>(m);
        return check_launch();
    }
    if (m.n_linear == 2 && m.dims[0] <= kSmMaxIn && m.dims[1] <= 128 && m.dims[2] <= kSmMaxOut) {
        // narrow generator (the README flow's S2 sub-pdf: 4 -> 128 -> 10): thread per row, see csrc/mlp_kernels.cuh
        static const bool off = [] { const char* e = getenv("JF_MLP_SMALL"); return e != nullptr && atoi(e) == 0; }();
        if (!off) {
            int dev = 0, sms = 148;
            JF_CUDA_OK(cudaGetDevice(&dev));
            JF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            const int64_t want = (B + 127) / 128;
            const unsigned grid = (unsigned)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
            auto go = [&](auto kern, int IN, int OUT) -> int {
                const size_t smem = (size_t)m.dims[1] * (IN + OUT + 1) * sizeof(T);
                JF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<grid, 128, smem, st>>>(m);
                return check_launch();
            };
            const int in = m.dims[0], out_n = m.dims[2];
            if (in <= 4 && out_n <= 4) return go(mlp_small_kernel<T, 4, 4>, 4, 4);
            if (in <= 4 && out_n <= 10) return go(mlp_small_kernel<T, 4, 10>, 4, 10);
            if (in <= 8 && out_n <= 10) return go(mlp_small_kernel<T, 8, 10>, 8, 10);
            if (in <= 8) return go(mlp_small_kernel<T, 8, 16>, 8, 16);
            if (out_n <= 10) return go(mlp_small_kernel<T, 16, 10>, 16, 10);
            return go(mlp_small_kernel<T, 16, 16>, 16, 16);
        }
    }
    if (!accumulate && ws != nullptr && i8_eligible(desc, sizeof(T) == 8 ? JF_F64 : JF_F32) &&
        ws_bytes >= i8_ws_bytes(desc->dims[2], sizeof(T) == 8 ? JF_F64 : JF_F32))
        return launch_mlp2_i8(m, ws, prepared, st);
    if (sizeof(T) == 8 && !accumulate) {
        const int rc = try_mlp2_dmma(m, st);
        if (rc != JF_ERR_UNSUPPORTED) return rc;
    }
    auto need = [&](int tm) { return (size_t)(2 * tm * m.lda + kMlpKC * kMlpLDW) * sizeof(T); };
    const size_t cap = 200 * 1024;
    if (need(64) <= cap) return launch_mlp<T, 64>(m, need(64), st);
    if (need(32) <= cap) return launch_mlp<T, 32>(m, need(32), st);
    if (need(16) <= cap) return launch_mlp<T, 16>(m, need(16), st);
    return JF_ERR_UNSUPPORTED;
}

extern "C" int jf_mlp_forward(const JfMlpDesc* desc, int dtype, const void* const* seg_ptrs, const int64_t* seg_ld,
                              const void* const* weights, const void* const* biases, void* out,
                              int64_t out_stride_param, int64_t out_stride_row, int64_t B, void* stream) {
    if (desc == nullptr || out == nullptr || seg_ptrs == nullptr || seg_ld == nullptr) return JF_ERR_BAD_ARG;
    if (desc->n_linear < 1 || desc->n_linear > JF_MAX_MLP_LINEAR) return JF_ERR_BAD_DESC;
    if (desc->n_segments < 1 || desc->n_segments > JF_MAX_MLP_SEGMENTS) return JF_ERR_BAD_DESC;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == JF_F64)
        return mlp_forward_t<double>(desc, seg_ptrs, seg_ld, weights, biases, out, out_stride_param, out_stride_row, B, st);
    if (dtype == JF_F32)
        return mlp_forward_t<float>(desc, seg_ptrs, seg_ld, weights, biases, out, out_stride_param, out_stride_row, B, st);
    return JF_ERR_BAD_ARG;
}

extern "C" int jf_mlp_forward_acc(const JfMlpDesc* desc, int dtype, const void* const* seg_ptrs, const int64_t* seg_ld,
                                  const void* const* weights, const void* const* biases, void* out,
                                  int64_t out_stride_param, int64_t out_stride_row, int64_t B, int accumulate,
                                  void* stream) {
    if (desc == nullptr || out == nullptr || seg_ptrs == nullptr || seg_ld == nullptr) return JF_ERR_BAD_ARG;
    if (desc->n_linear < 1 || desc->n_linear > JF_MAX_MLP_LINEAR) return JF_ERR_BAD_DESC;
    if (desc->n_segments < 1 || desc->n_segments > JF_MAX_MLP_SEGMENTS) return JF_ERR_BAD_DESC;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == JF_F64)
        return mlp_forward_t<double>(desc, seg_ptrs, seg_ld, weights, biases, out, out_stride_param, out_stride_row, B, st,
                                     nullptr, 0, 0, accumulate ? 1 : 0);
    if (dtype == JF_F32)
        return mlp_forward_t<float>(desc, seg_ptrs, seg_ld, weights, biases, out, out_stride_param, out_stride_row, B, st,
                                    nullptr, 0, 0, accumulate ? 1 : 0);
    return JF_ERR_BAD_ARG;
}

extern "C" int64_t jf_mlp_workspace_bytes(const JfMlpDesc* desc, int dtype) {
    if (desc == nullptr) return -1;
    return i8_eligible(desc, dtype) ? i8_ws_bytes(desc->dims[2], dtype) : 0;
}

extern "C" int jf_mlp_forward_ws(const JfMlpDesc* desc, int dtype, const void* const* seg_ptrs, const int64_t* seg_ld,
                                 const void* const* weights, const void* const* biases, void* out,
                                 int64_t out_stride_param, int64_t out_stride_row, int64_t B, void* workspace,
                                 int64_t workspace_bytes, int prepared, void* stream) {
    if (desc == nullptr || out == nullptr || seg_ptrs == nullptr || seg_ld == nullptr) return JF_ERR_BAD_ARG;
    if (desc->n_linear < 1 || desc->n_linear > JF_MAX_MLP_LINEAR) return JF_ERR_BAD_DESC;
    if (desc->n_segments < 1 || desc->n_segments > JF_MAX_MLP_SEGMENTS) return JF_ERR_BAD_DESC;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == JF_F64)
        return mlp_forward_t<double>(desc, seg_ptrs, seg_ld, weights, biases, out, out_stride_param, out_stride_row, B, st,
                                     workspace, workspace_bytes, prepared);
    if (dtype == JF_F32)
        return mlp_forward_t<float>(desc, seg_ptrs, seg_ld, weights, biases, out, out_stride_param, out_stride_row, B, st,
                                    workspace, workspace_bytes, prepared);
    return JF_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------------------------------------
// jf_mlp_backward: gradient of the parameter generator on the tensor cores (csrc/mlp_bwd.cuh)
// ---------------------------------------------------------------------------------------------------------------------
static bool mlp_bwd_eligible(const JfMlpDesc* md, int dtype) {
    return md != nullptr && dtype == JF_F32 && md->n_linear == 2 && md->dims[1] == 128 && md->dims[0] >= 1 &&
           md->dims[0] <= 96 && md->dims[2] >= 1;
}

extern "C" int64_t jf_mlp_backward_workspace_bytes(const JfMlpDesc* desc, int dtype, int64_t B) {
    if (!mlp_bwd_eligible(desc, dtype) || B < 0) return -1;
    return mlp_bwd_workspace_bytes(desc->dims[2], B);
}

extern "C" int jf_mlp_backward(const JfMlpDesc* desc, int dtype, const void* inp, int64_t ld_inp,
                               const void* const* weights, const void* const* biases, const void* grad_out,
                               int64_t go_stride_param, int64_t go_stride_row, const void* row_scale, void* grad_w1,
                               void* grad_b1, void* grad_w2,
                               void* grad_b2, void* grad_inp, int64_t ld_ginp, int64_t B, void* workspace,
                               int64_t workspace_bytes, void* stream) {
    if (desc == nullptr || inp == nullptr || weights == nullptr || biases == nullptr || grad_out == nullptr ||
        grad_w1 == nullptr || grad_b1 == nullptr || grad_w2 == nullptr || grad_b2 == nullptr)
        return JF_ERR_BAD_ARG;
    if (!mlp_bwd_eligible(desc, dtype)) return JF_ERR_UNSUPPORTED;
    if (go_stride_row != 1 || (go_stride_param & 3) != 0 || go_stride_param < B) return JF_ERR_BAD_ARG;   // param-major, 16-byte rows
    if ((reinterpret_cast<uintptr_t>(grad_out) & 15) != 0) return JF_ERR_BAD_ARG;
    if (B < 0) return JF_ERR_BAD_ARG;
    if (B == 0) return JF_OK;
    if (workspace == nullptr || workspace_bytes < mlp_bwd_workspace_bytes(desc->dims[2], B)) return JF_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return JF_ERR_WORKSPACE;
    BwArgs a;
    memset(&a, 0, sizeof(a));
    a.G = (const float*)grad_out; a.ldg = go_stride_param;
    a.row_scale = (const float*)row_scale;
    a.P = desc->dims[2]; a.B = B;
    a.x = (const float*)inp; a.ldx = ld_inp; a.in = desc->dims[0];
    a.W1 = (const float*)weights[0]; a.b1 = (const float*)biases[0]; a.W2 = (const float*)weights[1];
    if (a.W1 == nullptr || a.b1 == nullptr || a.W2 == nullptr) return JF_ERR_BAD_ARG;
    a.dW1 = (float*)grad_w1; a.db1 = (float*)grad_b1; a.dW2 = (float*)grad_w2; a.db2 = (float*)grad_b2;
    a.dx = (float*)grad_inp; a.lddx = ld_ginp;
    int rc = launch_mlp_bwd(a, workspace, (cudaStream_t)stream);
    g_launches.fetch_add(kMlpBwdLaunches - 1, std::memory_order_relaxed);
    if (rc != JF_OK) return rc;
    return check_launch();
}

// ---------------------------------------------------------------------------------------------------------------------
// jf_subpdf_apply_generated: parameter generator + layer chain of one conditional Euclidean sub-pdf in ONE kernel
// (csrc/gf_fused.cuh): the per-row parameters never exist in HBM
// ---------------------------------------------------------------------------------------------------------------------
static bool fused_eligible(const JfSubPdfDesc* sp, const JfMlpDesc* md, int dtype) {
    static const bool off = [] { const char* e = getenv("JF_FUSED"); return e != nullptr && atoi(e) == 0; }();
    if (off || dtype != JF_F64 || sp == nullptr || md == nullptr) return false;
    if (sp->manifold != 'e' || sp->dim < 1 || sp->dim > kFuMaxD) return false;
    if (sp->n_layers < 1 || sp->n_layers > JF_MAX_LAYERS) return false;
    if (md->n_linear != 2 || md->dims[1] != kI8H || md->dims[0] < 1 || md->dims[0] > kI8MaxKin) return false;
    if (md->dims[2] < sp->n_params) return false;
    int off_p = 0;
    for (int l = 0; l < sp->n_layers; ++l) {
        const JfLayerDesc& L = sp->layers[l];
        if (L.kind != JF_LAYER_GF || !gf_layer_is_default(L) || L.dim != sp->dim) return false;
        if (L.K != kFuK || L.norm_mode != JF_NORM_REGULATED) return false;
        if (L.hh_iter < 0 || L.hh_iter > kFuMaxHH) return false;
        if (L.inv_type < 0 || L.inv_type > 3) return false;
        if (!(L.w_min > 0) || !(L.w_max > 0)) return false;
        if (L.param_offset != off_p) return false;
        if ((L.has_offset ? sp->dim : 0) + L.hh_iter * sp->dim + 3 * L.K * sp->dim != L.n_params) return false;
        off_p += L.n_params;
    }
    return off_p == sp->n_params;
}

// eligible AND the kernel of this direction fits into the device's shared memory (sampling keeps the kernels of every
// (row, dimension) in shared-memory slots, which limits the generator's input width)
static bool fused_usable(const JfSubPdfDesc* sp, const JfMlpDesc* md, int dtype, int direction) {
    if (!fused_eligible(sp, md, dtype)) return false;
    static const int smem_max = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
        return v;
    }();
    return fused_fits(direction, md->dims[0], smem_max);
}

extern "C" int64_t jf_subpdf_generated_workspace_bytes(const JfSubPdfDesc* desc, const JfMlpDesc* mlp, int dtype) {
    if (!fused_eligible(desc, mlp, dtype)) return -1;
    return fused_prep_bytes(desc->n_layers);
}

extern "C" int jf_subpdf_apply_generated(const JfSubPdfDesc* desc, const JfMlpDesc* mlp, int dtype, int direction,
                                         const void* const* seg_ptrs, const int64_t* seg_ld,
                                         const void* const* weights, const void* const* biases, const void* in,
                                         int64_t ld_in, const void* logdet_in, void* logdet_out, const void* logbase_in,
                                         void* logbase_out, void* out, int64_t ld_out, int64_t B, void* workspace,
                                         int64_t workspace_bytes, int prepared, int64_t* status, void* stream) {
    if (desc == nullptr || mlp == nullptr || in == nullptr || out == nullptr || seg_ptrs == nullptr || seg_ld == nullptr ||
        weights == nullptr || biases == nullptr)
        return JF_ERR_BAD_ARG;
    if (direction != JF_DIR_LOGPDF && direction != JF_DIR_SAMPLE) return JF_ERR_BAD_ARG;
    if (!fused_usable(desc, mlp, dtype, direction)) return JF_ERR_UNSUPPORTED;
    if (mlp->n_segments < 1 || mlp->n_segments > JF_MAX_MLP_SEGMENTS) return JF_ERR_BAD_DESC;
    if (workspace == nullptr || workspace_bytes < fused_prep_bytes(desc->n_layers)) return JF_ERR_WORKSPACE;
    if (B < 0) return JF_ERR_BAD_ARG;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    FuArgs a;
    memset(&a, 0, sizeof(a));
    a.m.n_linear = 2;
    int in_sum = 0;
    for (int l = 0; l <= 2; ++l) a.m.dims[l] = mlp->dims[l];
    a.m.n_segments = mlp->n_segments;
    for (int s = 0; s < mlp->n_segments; ++s) {
        if (seg_ptrs[s] == nullptr || mlp->seg_cols[s] < 1) return JF_ERR_BAD_ARG;
        a.m.seg_cols[s] = mlp->seg_cols[s];
        a.m.seg_ptr[s] = (const double*)seg_ptrs[s];
        a.m.seg_ld[s] = seg_ld[s];
        in_sum += mlp->seg_cols[s];
    }
    if (in_sum != mlp->dims[0]) return JF_ERR_BAD_DESC;
    for (int l = 0; l < 2; ++l) {
        a.m.wt[l] = (const double*)weights[l];
        a.m.bias[l] = (const double*)biases[l];
        if (a.m.wt[l] == nullptr || a.m.bias[l] == nullptr) return JF_ERR_BAD_ARG;
    }
    a.m.B = B;
    a.n_layers = desc->n_layers;
    a.d = desc->dim;
    for (int l = 0; l < desc->n_layers; ++l) {
        const JfLayerDesc& L = desc->layers[l];
        FuLayerC& c = a.layers[l];
        c.inv_type = L.inv_type; c.has_offset = L.has_offset; c.hh_iter = L.hh_iter; c.raw_off = L.param_offset;
        c.w_min = L.w_min; c.inv_w_max = 1.0 / L.w_max; c.n_min = L.n_min; c.n_max = L.n_max;
    }
    a.in = (const double*)in; a.ld_in = ld_in;
    a.out = (double*)out; a.ld_out = ld_out;
    a.logdet_in = (const double*)logdet_in; a.logdet_out = (double*)logdet_out;
    a.logbase_in = (const double*)logbase_in; a.logbase_out = (double*)logbase_out;
    a.status = status;
    int rc = launch_fused_prep(a, a.m.wt[1], a.m.bias[1], direction, workspace, !prepared, st);
    if (rc != JF_OK) return rc;
    if (!prepared) { rc = check_launch(); if (rc != JF_OK) return rc; }
    rc = launch_fused(a, direction, st);
    if (rc != JF_OK) return rc;
    return check_launch();
}

// ---------------------------------------------------------------------------------------------------------------------
// whole-pdf orchestration (chunked; every launch on the caller's stream)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void combine_kernel(const T* a, const T* b, T sign_b, T* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + sign_b * b[i];
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static inline size_t esize(int dtype) { return dtype == JF_F64 ? 8 : 4; }

struct WsLayout {
    int64_t params, emb[JF_MAX_SUBPDFS], mlp[JF_MAX_SUBPDFS], mlp_bytes[JF_MAX_SUBPDFS], logdet, logbase, scratch, total;
};

static int ws_layout(const JfPdfDesc* d, int64_t chunk, WsLayout& w) {
    if (d == nullptr || d->abi_version != JF_ABI_VERSION) return JF_ERR_BAD_DESC;
    if (d->n_sub < 1 || d->n_sub > JF_MAX_SUBPDFS || chunk < 1) return JF_ERR_BAD_DESC;
    const int64_t es = (int64_t)esize(d->dtype);
    int64_t off = 0;
    int pmax = 0;
    for (int k = 0; k < d->n_sub; ++k)
        if (d->has_mlp[k] && d->sub[k].n_params > pmax) pmax = d->sub[k].n_params;
    w.params = off; off = align_up(off + (int64_t)pmax * chunk * es, 256);
    for (int k = 0; k < d->n_sub; ++k) {
        w.emb[k] = off;
        if (d->sub[k].manifold == 's') off = align_up(off + (int64_t)d->emb_dim[k] * chunk * es, 256);
    }
    for (int k = 0; k < d->n_sub; ++k) {
        w.mlp[k] = off;
        w.mlp_bytes[k] = 0;
        if (d->has_mlp[k]) {
            JfMlpDesc md = d->mlp[k];           // jf_mlp_workspace_bytes only looks at n_linear / dims
            int64_t nb = jf_mlp_workspace_bytes(&md, d->dtype);
            const int64_t nf = jf_subpdf_generated_workspace_bytes(&d->sub[k], &md, d->dtype);
            if (nf > nb) nb = nf;
            if (nb > 0) { w.mlp_bytes[k] = nb; off = align_up(off + nb, 256); }
        }
    }
    w.logdet = off; off = align_up(off + chunk * es, 256);
    w.logbase = off; off = align_up(off + chunk * es, 256);
    const int maxcols = d->total_base_dim > d->total_target_dim ? d->total_base_dim : d->total_target_dim;
    w.scratch = off; off = align_up(off + (int64_t)maxcols * chunk * es, 256);
    w.total = off;
    return JF_OK;
}

extern "C" int64_t jf_pdf_workspace_bytes(const JfPdfDesc* desc, int64_t chunk_rows) {
    WsLayout w;
    if (ws_layout(desc, chunk_rows, w) != JF_OK) return -1;
    return w.total;
}

// direction-generic chunk loop.  `src` = x (logpdf) or z (sample); `dst` = base (logpdf) or x (sample).
static int pdf_run(const JfPdfDesc* d, const JfPdfParams* P, int direction, const void* src, int64_t ld_src,
                   const void* cond, int64_t ldc, void* dst, int64_t ld_dst, void* logp, void* logp_base, int64_t B,
                   void* workspace, int64_t ws_bytes, int64_t chunk, int64_t* status, cudaStream_t st) {
    WsLayout w;
    int rc = ws_layout(d, chunk, w);
    if (rc != JF_OK) return rc;
    if (B == 0) return JF_OK;
    if (B < 0) return JF_ERR_BAD_ARG;
    if (P == nullptr || src == nullptr || (d->cond_dim > 0 && cond == nullptr)) return JF_ERR_BAD_ARG;
    if (workspace == nullptr || ws_bytes < w.total) return JF_ERR_WORKSPACE;
    const int64_t es = (int64_t)esize(d->dtype);
    char* ws = (char*)workspace;
    const bool logpdf = direction == JF_DIR_LOGPDF;
    for (int64_t r0 = 0; r0 < B; r0 += chunk) {
        const int64_t n = (B - r0 < chunk) ? (B - r0) : chunk;
        const char* src_c = (const char*)src + r0 * ld_src * es;
        char* dst_c;
        int64_t ld_dst_c;
        if (dst != nullptr) { dst_c = (char*)dst + r0 * ld_dst * es; ld_dst_c = ld_dst; }
        else { dst_c = ws + w.scratch; ld_dst_c = logpdf ? d->total_base_dim : d->total_target_dim; }
        // target-space coordinates of this chunk (input for logpdf, output for sampling): source of 'e' embeddings
        const char* tgt_c = logpdf ? src_c : dst_c;
        const int64_t ld_tgt = logpdf ? ld_src : ld_dst_c;
        void* logdet = ws + w.logdet;
        void* logbase = ws + w.logbase;
        for (int k = 0; k < d->n_sub; ++k) {
            const JfSubPdfDesc* sp = &d->sub[k];
            const void* params = P->shared[k];
            int64_t sj = 1, sr = 0;
            if (d->has_mlp[k]) {
                JfMlpDesc md = d->mlp[k];
                const void* seg_ptr[JF_MAX_MLP_SEGMENTS];
                int64_t seg_ld[JF_MAX_MLP_SEGMENTS];
                int ns = 0;
                if (d->cond_dim > 0) {
                    md.seg_cols[ns] = d->cond_dim;
                    seg_ptr[ns] = (const char*)cond + r0 * ldc * es;
                    seg_ld[ns] = ldc;
                    ++ns;
                }
                for (int j = 0; j < k; ++j) {
                    if (ns >= JF_MAX_MLP_SEGMENTS) return JF_ERR_UNSUPPORTED;
                    md.seg_cols[ns] = d->emb_dim[j];
                    if (d->sub[j].manifold != 's') {
                        seg_ptr[ns] = tgt_c + (int64_t)d->target_col[j] * es;
                        seg_ld[ns] = ld_tgt;
                    } else {
                        seg_ptr[ns] = ws + w.emb[j];
                        seg_ld[ns] = d->emb_dim[j];
                    }
                    ++ns;
                }
                md.n_segments = ns;
                if (w.mlp_bytes[k] > 0 && fused_usable(sp, &md, d->dtype, direction)) {
                    // generator + layer chain in one kernel: the parameter block never touches HBM
                    const int in_col_f = logpdf ? d->target_col[k] : d->base_col[k];
                    const int out_col_f = logpdf ? d->base_col[k] : d->target_col[k];
                    rc = jf_subpdf_apply_generated(sp, &md, d->dtype, direction, seg_ptr, seg_ld, P->weights[k], P->biases[k],
                                                   src_c + (int64_t)in_col_f * es, ld_src, k == 0 ? nullptr : logdet, logdet,
                                                   k == 0 ? nullptr : logbase, logbase, dst_c + (int64_t)out_col_f * es,
                                                   ld_dst_c, n, ws + w.mlp[k], w.mlp_bytes[k], r0 > 0 ? 1 : 0, status, st);
                    if (rc != JF_OK) return rc;
                    continue;
                }
                // the sliced last-layer weights are prepared by the first chunk and reused by the later ones
                rc = jf_mlp_forward_ws(&md, d->dtype, seg_ptr, seg_ld, P->weights[k], P->biases[k], ws + w.params, chunk,
                                       1, n, w.mlp_bytes[k] > 0 ? ws + w.mlp[k] : nullptr, w.mlp_bytes[k], r0 > 0 ? 1 : 0,
                                       st);
                if (rc != JF_OK) return rc;
                params = ws + w.params;
                sj = chunk;
                sr = 1;
            } else if (sp->n_params > 0 && params == nullptr) {
                return JF_ERR_BAD_ARG;
            }
            const int in_col = logpdf ? d->target_col[k] : d->base_col[k];
            const int out_col = logpdf ? d->base_col[k] : d->target_col[k];
            bool need_emb = false;
            for (int j = k + 1; j < d->n_sub; ++j) need_emb = need_emb || d->has_mlp[j];
            void* emb = (need_emb && sp->manifold == 's') ? (void*)(ws + w.emb[k]) : nullptr;
            rc = jf_subpdf_apply(sp, d->dtype, direction, src_c + (int64_t)in_col * es, ld_src, params, sj, sr,
                                 k == 0 ? nullptr : logdet, logdet, k == 0 ? nullptr : logbase, logbase,
                                 dst_c + (int64_t)out_col * es, ld_dst_c, emb, d->emb_dim[k], n, status, st);
            if (rc != JF_OK) return rc;
        }
        // log p = log N(base) + logdet (log_pdf direction) / log N(z) - logdet (sampling direction)
        if (logp != nullptr) {
            const unsigned blocks = (unsigned)((n + 255) / 256);
            if (d->dtype == JF_F64)
                combine_kernel<double><<<blocks, 256, 0, st>>>((const double*)logbase, (const double*)logdet,
                                                               logpdf ? 1.0 : -1.0, (double*)logp + r0, n);
            else
                combine_kernel<float><<<blocks, 256, 0, st>>>((const float*)logbase, (const float*)logdet,
                                                              logpdf ? 1.f : -1.f, (float*)logp + r0, n);
            rc = check_launch();
            if (rc != JF_OK) return rc;
        }
        if (logp_base != nullptr)
            JF_CUDA_OK(cudaMemcpyAsync((char*)logp_base + r0 * es, logbase, n * es, cudaMemcpyDeviceToDevice, st));
    }
    return JF_OK;
}

extern "C" int jf_pdf_logpdf(const JfPdfDesc* desc, const JfPdfParams* params, const void* x, int64_t ldx,
                             const void* cond, int64_t ldc, void* logp, void* logp_base, void* base, int64_t ld_base,
                             int64_t B, void* workspace, int64_t workspace_bytes, int64_t chunk_rows, int64_t* status,
                             void* stream) {
    return pdf_run(desc, params, JF_DIR_LOGPDF, x, ldx, cond, ldc, base, ld_base, logp, logp_base, B, workspace,
                   workspace_bytes, chunk_rows, status, (cudaStream_t)stream);
}

extern "C" int jf_pdf_sample(const JfPdfDesc* desc, const JfPdfParams* params, const void* z, int64_t ldz,
                             const void* cond, int64_t ldc, void* x, int64_t ldx, void* logp, void* logp_base, int64_t B,
                             void* workspace, int64_t workspace_bytes, int64_t chunk_rows, int64_t* status, void* stream) {
    return pdf_run(desc, params, JF_DIR_SAMPLE, z, ldz, cond, ldc, x, ldx, logp, logp_base, B, workspace,
                   workspace_bytes, chunk_rows, status, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// host-buffer entries: H2D -> kernels -> D2H pipelined over two streams, two buffer sets
// ---------------------------------------------------------------------------------------------------------------------
struct HostWs {
    int64_t src, cond, dst, logp, logp_base, inner, per_set, total;
};

static int host_ws_layout(const JfPdfDesc* d, int64_t chunk, HostWs& h) {
    WsLayout w;
    int rc = ws_layout(d, chunk, w);
    if (rc != JF_OK) return rc;
    const int64_t es = (int64_t)esize(d->dtype);
    const int maxcols = d->total_base_dim > d->total_target_dim ? d->total_base_dim : d->total_target_dim;
    int64_t off = 0;
    h.src = off; off = align_up(off + (int64_t)maxcols * chunk * es, 256);
    h.cond = off; off = align_up(off + (int64_t)(d->cond_dim > 0 ? d->cond_dim : 1) * chunk * es, 256);
    h.dst = off; off = align_up(off + (int64_t)maxcols * chunk * es, 256);
    h.logp = off; off = align_up(off + chunk * es, 256);
    h.logp_base = off; off = align_up(off + chunk * es, 256);
    h.inner = off; off = align_up(off + w.total, 256);
    h.per_set = off;
    h.total = 2 * off;
    return JF_OK;
}

extern "C" int64_t jf_pdf_host_workspace_bytes(const JfPdfDesc* desc, int64_t chunk_rows) {
    HostWs h;
    if (host_ws_layout(desc, chunk_rows, h) != JF_OK) return -1;
    return h.total;
}

// strided [n, cols] copy; collapses to ONE linear copy when both sides are dense (a 2-D copy of millions of 80-byte
// rows is an order of magnitude slower on the copy engines)
static cudaError_t copy_rows(void* dst, int64_t ld_dst, const void* src, int64_t ld_src, int64_t cols, int64_t n,
                             int64_t es, cudaMemcpyKind kind, cudaStream_t s) {
    if (ld_dst == cols && ld_src == cols)
        return cudaMemcpyAsync(dst, src, (size_t)(n * cols * es), kind, s);
    return cudaMemcpy2DAsync(dst, (size_t)(ld_dst * es), src, (size_t)(ld_src * es), (size_t)(cols * es), (size_t)n, kind, s);
}

static int pdf_run_host(const JfPdfDesc* d, const JfPdfParams* P, int direction, const void* src_h, int64_t ld_src,
                        const void* cond_h, int64_t ldc, void* dst_h, int64_t ld_dst, void* logp_h, void* logp_base_h,
                        int64_t B, void* workspace, int64_t ws_bytes, int64_t chunk, int64_t* status) {
    HostWs h;
    int rc = host_ws_layout(d, chunk, h);
    if (rc != JF_OK) return rc;
    if (workspace == nullptr || ws_bytes < h.total) return JF_ERR_WORKSPACE;
    if (src_h == nullptr || (d->cond_dim > 0 && cond_h == nullptr)) return JF_ERR_BAD_ARG;
    WsLayout w;
    ws_layout(d, chunk, w);
    const int64_t es = (int64_t)esize(d->dtype);
    const bool logpdf = direction == JF_DIR_LOGPDF;
    const int src_cols = logpdf ? d->total_target_dim : d->total_base_dim;
    const int dst_cols = logpdf ? d->total_base_dim : d->total_target_dim;
    // two copy/compute streams + two events per (host thread, device), created on first use and kept (creating and
    // destroying them cost ~50 us per call).  They are non-blocking streams: the CALLER must have finished preparing the
    // parameter vector / workspace / status words (jammy_flows_b200.engine synchronises torch's current stream first).
    struct HostStreams { cudaStream_t st[2]; cudaEvent_t done[2]; bool ready; };
    static thread_local HostStreams cache[64] = {};
    int dev = 0;
    JF_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return JF_ERR_BAD_ARG;
    HostStreams& hs = cache[dev];
    if (!hs.ready) {
        JF_CUDA_OK(cudaStreamCreateWithFlags(&hs.st[0], cudaStreamNonBlocking));
        JF_CUDA_OK(cudaStreamCreateWithFlags(&hs.st[1], cudaStreamNonBlocking));
        JF_CUDA_OK(cudaEventCreateWithFlags(&hs.done[0], cudaEventDisableTiming));
        JF_CUDA_OK(cudaEventCreateWithFlags(&hs.done[1], cudaEventDisableTiming));
        hs.ready = true;
    }
    cudaStream_t* st = hs.st;
    cudaEvent_t* done = hs.done;
    // the two streams exist to overlap COPIES with kernels; the kernels of consecutive chunks are chained by events so
    // that they never share the SMs (a persistent MLP CTA next to layer-kernel CTAs of the other chunk slows both)
    static const bool chain_compute = [] { const char* e = getenv("JF_HOST_OVERLAP_COMPUTE"); return !(e && atoi(e) == 1); }();
    int64_t ci = 0;
    for (int64_t r0 = 0; r0 < B && rc == JF_OK; r0 += chunk, ++ci) {
        const int64_t n = (B - r0 < chunk) ? (B - r0) : chunk;
        const int b = (int)(ci & 1);
        char* set = (char*)workspace + (int64_t)b * h.per_set;
        cudaStream_t s = st[b];
        cudaError_t e = copy_rows(set + h.src, src_cols, (const char*)src_h + r0 * ld_src * es, ld_src, src_cols, n, es,
                                  cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && d->cond_dim > 0)
            e = copy_rows(set + h.cond, d->cond_dim, (const char*)cond_h + r0 * ldc * es, ldc, d->cond_dim, n, es,
                          cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { rc = (int)e; break; }
        if (chain_compute && ci > 0) {
            e = cudaStreamWaitEvent(s, done[1 - b], 0);
            if (e != cudaSuccess) { rc = (int)e; break; }
        }
        rc = pdf_run(d, P, direction, set + h.src, src_cols, d->cond_dim > 0 ? set + h.cond : nullptr, d->cond_dim,
                     set + h.dst, dst_cols, set + h.logp, logp_base_h ? set + h.logp_base : nullptr, n, set + h.inner,
                     w.total, chunk, status, s);
        if (rc != JF_OK) break;
        e = cudaEventRecord(done[b], s);
        if (e != cudaSuccess) { rc = (int)e; break; }
        if (dst_h != nullptr)
            e = copy_rows((char*)dst_h + r0 * ld_dst * es, ld_dst, set + h.dst, dst_cols, dst_cols, n, es,
                          cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && logp_h != nullptr)
            e = cudaMemcpyAsync((char*)logp_h + r0 * es, set + h.logp, n * es, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && logp_base_h != nullptr)
            e = cudaMemcpyAsync((char*)logp_base_h + r0 * es, set + h.logp_base, n * es, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) { rc = (int)e; break; }
    }
    cudaError_t e0 = cudaStreamSynchronize(st[0]);
    cudaError_t e1 = cudaStreamSynchronize(st[1]);
    if (rc != JF_OK) return rc;
    if (e0 != cudaSuccess) return (int)e0;
    if (e1 != cudaSuccess) return (int)e1;
    return JF_OK;
}

extern "C" int jf_pdf_logpdf_host(const JfPdfDesc* desc, const JfPdfParams* params, const void* x_host, int64_t ldx,
                                  const void* cond_host, int64_t ldc, void* logp_host, void* logp_base_host,
                                  void* base_host, int64_t ld_base, int64_t B, void* workspace, int64_t workspace_bytes,
                                  int64_t chunk_rows, int64_t* status) {
    return pdf_run_host(desc, params, JF_DIR_LOGPDF, x_host, ldx, cond_host, ldc, base_host, ld_base, logp_host,
                        logp_base_host, B, workspace, workspace_bytes, chunk_rows, status);
}

extern "C" int jf_pdf_sample_host(const JfPdfDesc* desc, const JfPdfParams* params, const void* z_host, int64_t ldz,
                                  const void* cond_host, int64_t ldc, void* x_host, int64_t ldx, void* logp_host,
                                  void* logp_base_host, int64_t B, void* workspace, int64_t workspace_bytes,
                                  int64_t chunk_rows, int64_t* status) {
    return pdf_run_host(desc, params, JF_DIR_SAMPLE, z_host, ldz, cond_host, ldc, x_host, ldx, logp_host, logp_base_host,
                        B, workspace, workspace_bytes, chunk_rows, status);
}

// ---------------------------------------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
static int transform_target(const JfPdfDesc* desc, int to_embedding, const void* in, int64_t ld_in, void* out,
                            int64_t ld_out, const void* logdet_in, void* logdet_out, int64_t B, cudaStream_t st) {
    ChartArgs<T> a;
    memset(&a, 0, sizeof(a));
    a.n_sub = desc->n_sub; a.to_embedding = to_embedding ? 1 : 0;
    int ci = 0, ce = 0;   // running intrinsic / embedded column
    for (int k = 0; k < desc->n_sub; ++k) {
        const JfSubPdfDesc& s = desc->sub[k];
        const bool sphere = s.manifold == 's';
        if (sphere && s.dim != 1 && s.dim != 2) return JF_ERR_UNSUPPORTED;
        a.kind[k] = sphere ? s.dim : 0;
        a.dim[k] = s.dim;
        a.in_col[k] = to_embedding ? ci : ce;
        a.out_col[k] = to_embedding ? ce : ci;
        ci += s.dim;
        ce += s.dim + (sphere ? 1 : 0);
    }
    a.in = (const T*)in; a.ld_in = ld_in; a.out = (T*)out; a.ld_out = ld_out;
    a.logdet_in = (const T*)logdet_in; a.logdet_out = (T*)logdet_out; a.B = B;
    const int64_t blocks = (B + 255) / 256;
    chart_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a);
    return check_launch();
}

extern "C" int jf_pdf_transform_target(const JfPdfDesc* desc, int to_embedding, const void* in, int64_t ld_in, void* out,
                                       int64_t ld_out, const void* logdet_in, void* logdet_out, int64_t B, void* stream) {
    if (desc == nullptr || desc->abi_version != JF_ABI_VERSION) return JF_ERR_BAD_DESC;
    if (desc->n_sub < 1 || desc->n_sub > JF_MAX_SUBPDFS) return JF_ERR_BAD_DESC;
    if (B < 0 || (B > 0 && (in == nullptr || out == nullptr))) return JF_ERR_BAD_ARG;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (desc->dtype == JF_F64) return transform_target<double>(desc, to_embedding, in, ld_in, out, ld_out, logdet_in, logdet_out, B, st);
    if (desc->dtype == JF_F32) return transform_target<float>(desc, to_embedding, in, ld_in, out, ld_out, logdet_in, logdet_out, B, st);
    return JF_ERR_BAD_ARG;
}

extern "C" int jf_row_logmeanexp(int dtype, const void* in, int64_t rows, int64_t cols, void* out, void* stream) {
    if (rows < 0 || cols < 1 || (rows > 0 && (in == nullptr || out == nullptr))) return JF_ERR_BAD_ARG;
    if (rows == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t blocks = (rows * 32 + 255) / 256;
    if (dtype == JF_F64) row_logmeanexp_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((const double*)in, rows, cols, (double*)out);
    else if (dtype == JF_F32) row_logmeanexp_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)in, rows, cols, (float*)out);
    else return JF_ERR_BAD_ARG;
    return check_launch();
}

extern "C" int jf_normal_rows(int dtype, uint64_t seed, uint64_t first_row, int64_t B, int32_t dim, void* out,
                              int64_t ld_out, void* stream) {
    if (B < 0 || dim < 1 || ld_out < dim || (B > 0 && out == nullptr)) return JF_ERR_BAD_ARG;
    if (B == 0) return JF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = B * ((dim + 1) / 2);
    const int64_t blocks = (n + 255) / 256;
    if (dtype == JF_F64) normal_rows_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(seed, first_row, B, dim, (double*)out, ld_out);
    else if (dtype == JF_F32) normal_rows_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(seed, first_row, B, dim, (float*)out, ld_out);
    else return JF_ERR_BAD_ARG;
    return check_launch();
}

template <typename T>
static int rowwise_linear_t(const void* params, int64_t ld_p, int64_t off_w, int64_t off_b, const void* in, int64_t ld_in,
                            int n_in, int n_out, int act, int accumulate, void* out, int64_t so_p, int64_t so_r, int64_t B,
                            cudaStream_t st) {
    const int n_in_pad = (n_in + 1) & ~1;
    const size_t smem = (size_t)ROWWISE_WARPS * (n_in_pad + (n_in >= 32 ? 0 : ROWWISE_TILE)) * sizeof(T);
    if (smem > 200 * 1024) return JF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        JF_CUDA_OK(cudaFuncSetAttribute(rowwise_linear_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    JF_CUDA_OK(cudaGetDevice(&dev));
    JF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t blocks = (B + ROWWISE_WARPS - 1) / ROWWISE_WARPS;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    rowwise_linear_kernel<T><<<(unsigned)blocks, ROWWISE_WARPS * 32, smem, st>>>(
        (const T*)params, ld_p, off_w, off_b, (const T*)in, ld_in, n_in, n_in_pad, n_out, act, accumulate, (T*)out, so_p,
        so_r, B);
    return check_launch();
}

extern "C" int jf_rowwise_linear(int dtype, const void* params, int64_t ld_params, int64_t off_w, int64_t off_b,
                                 const void* in, int64_t ld_in, int32_t n_in, int32_t n_out, int act, int accumulate,
                                 void* out, int64_t out_stride_param, int64_t out_stride_row, int64_t B, void* stream) {
    if (B < 0 || n_in < 1 || n_out < 1 || off_w < 0 || ld_in < n_in) return JF_ERR_BAD_ARG;
    if (B == 0) return JF_OK;
    if (params == nullptr || in == nullptr || out == nullptr) return JF_ERR_BAD_ARG;
    if (off_w + (int64_t)n_in * n_out > ld_params || (off_b >= 0 && off_b + n_out > ld_params)) return JF_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == JF_F64)
        return rowwise_linear_t<double>(params, ld_params, off_w, off_b, in, ld_in, n_in, n_out, act ? 1 : 0,
                                        accumulate ? 1 : 0, out, out_stride_param, out_stride_row, B, st);
    if (dtype == JF_F32)
        return rowwise_linear_t<float>(params, ld_params, off_w, off_b, in, ld_in, n_in, n_out, act ? 1 : 0,
                                       accumulate ? 1 : 0, out, out_stride_param, out_stride_row, B, st);
    return JF_ERR_BAD_ARG;
}

extern "C" int jf_abi_version(void) { return JF_ABI_VERSION; }
extern "C" int64_t jf_launch_count(void) { return g_launches.load(); }
extern "C" int64_t jf_struct_size(int which) {
    switch (which) {
        case 0: return sizeof(JfLayerDesc);
        case 1: return sizeof(JfSubPdfDesc);
        case 2: return sizeof(JfMlpDesc);
        case 3: return sizeof(JfPdfDesc);
        case 4: return sizeof(JfPdfParams);
        case 5: return sizeof(JfSplineDesc);
        default: return -1;
    }
}

template <typename T>
__global__ void fma_probe_kernel(T* out, int iters, T a, T b) {
    T acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = T(threadIdx.x) * T(1e-3) + T(i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == T(12345.678)) out[0] = s;
}

extern "C" int jf_probe_fma_peak(int dtype, int iters, float* ms, double* fma_count, void* scratch, void* stream) {
    if (ms == nullptr || fma_count == nullptr || scratch == nullptr || iters < 1) return JF_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    JF_CUDA_OK(cudaGetDevice(&dev));
    JF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    void* buf = scratch;                  // 256 bytes of device memory from the caller: the library never allocates
    const int threads = 512, grid = sms * 4;
    cudaEvent_t e0, e1;
    JF_CUDA_OK(cudaEventCreate(&e0));
    JF_CUDA_OK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {   // first pass warms up
        JF_CUDA_OK(cudaEventRecord(e0, st));
        if (dtype == JF_F64) fma_probe_kernel<double><<<grid, threads, 0, st>>>((double*)buf, iters, 1.0000001, 1e-9);
        else fma_probe_kernel<float><<<grid, threads, 0, st>>>((float*)buf, iters, 1.0000001f, 1e-9f);
        check_launch();
        JF_CUDA_OK(cudaEventRecord(e1, st));
        JF_CUDA_OK(cudaEventSynchronize(e1));
    }
    JF_CUDA_OK(cudaEventElapsedTime(ms, e0, e1));
    *fma_count = (double)grid * threads * (double)iters * 8.0;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return JF_OK;
}
