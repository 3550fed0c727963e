// Sub-pdf kernels: all layers of one sub-pdf, fused, one thread per row.
#pragma once
#include "subpdf_args.cuh"
#include "gf.cuh"
#include "s2.cuh"
#include "chain1.cuh"

namespace jf {

template <typename T>
struct GfChainArgs {
    SubPdfArgs<T> a;
    GfLayerC<T> layers[JF_MAX_LAYERS];
};

// Euclidean sub-pdf: chain of "g" layers.
//   D_ > 0: compile-time dimension (row vector fully in registers); D_ == 0: run-time d <= JF_MAX_DIM.
//   num_kde is a run-time value (rolled loops over K).
// Dynamic shared memory: the processed table (shared parameters) or 3*Kmax*blockDim.x per-thread slots (per-row).
#ifndef JF_GF_MIN_BLOCKS
#define JF_GF_MIN_BLOCKS 3
#endif
template <typename T, int D_, int DIR>
__global__ void __launch_bounds__(256, JF_GF_MIN_BLOCKS) gf_chain_kernel(const __grid_constant__ GfChainArgs<T> g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* slots = tab;
    const SubPdfArgs<T>& a = g.a;
    constexpr int DM = D_ > 0 ? D_ : JF_MAX_DIM;
    const int d = D_ > 0 ? D_ : a.d;
    const bool shared_params = (a.sr == 0);

    if (shared_params) {
        // permanent parameters: regulate once per CTA into shared memory (broadcast reads afterwards)
        for (int l = 0; l < a.n_layers; ++l)
            if (g.layers[l].kind == 0) gf_build_table<T>(g.layers[l], a.params, tab, threadIdx.x, blockDim.x);
        __syncthreads();
    }

    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.B) return;

    T x[DM];
#pragma unroll
    for (int j = 0; j < d; ++j) x[j] = a.in[row * a.ld_in + j];
    T logdet = a.logdet_in ? a.logdet_in[row] : T(0);
    const T* prow = a.params + row * a.sr;

    T zsq = 0;   // sum of squares of the base-space coordinates
    if (DIR == JF_DIR_LOGPDF) {
        if (a.emb_out) {
#pragma unroll
            for (int j = 0; j < d; ++j) a.emb_out[row * a.ld_emb + j] = x[j];
        }
        for (int l = a.n_layers - 1; l >= 0; --l) {
            if (g.layers[l].kind == 1) {
                T tmp[DM];
#pragma unroll
                for (int j = 0; j < d; ++j) tmp[j] = x[j];
                mvn_layer_cold<T, DM>(true, tmp, &logdet, &g.layers[l], d, prow, a.sj);
#pragma unroll
                for (int j = 0; j < d; ++j) x[j] = tmp[j];
            }
            else gf_layer_logpdf<T, DM>(x, logdet, g.layers[l], d, g.layers[l].K, shared_params, tab, prow, a.sj, slots);
        }
#pragma unroll
        for (int j = 0; j < d; ++j) zsq = fma(x[j], x[j], zsq);
    } else {
#pragma unroll
        for (int j = 0; j < d; ++j) zsq = fma(x[j], x[j], zsq);
        int n_evals = 0, n_unconv = 0;
        for (int l = 0; l < a.n_layers; ++l) {
            if (g.layers[l].kind == 1) {
                T tmp[DM];
#pragma unroll
                for (int j = 0; j < d; ++j) tmp[j] = x[j];
                mvn_layer_cold<T, DM>(false, tmp, &logdet, &g.layers[l], d, prow, a.sj);
#pragma unroll
                for (int j = 0; j < d; ++j) x[j] = tmp[j];
            }
            else gf_layer_sample<T, DM>(x, logdet, g.layers[l], d, g.layers[l].K, shared_params, tab, prow, a.sj, slots,
                                        n_evals, n_unconv);
        }
        if (a.emb_out) {
#pragma unroll
            for (int j = 0; j < d; ++j) a.emb_out[row * a.ld_emb + j] = x[j];
        }
        if (n_unconv) status_add(a.status, JF_STATUS_UNCONVERGED, n_unconv);
        // evaluation counter: one atomic per warp
        unsigned m = __activemask();
        int tot = n_evals;
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(m, tot, o);
        if ((threadIdx.x & 31) == 0) status_add(a.status, JF_STATUS_ITERATIONS, tot);
    }
    bool bad = !finite_(logdet);
#pragma unroll
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

constexpr int kS2MaxSplines = 8;   // nested spline sub-flows of all "f" layers of one sub-pdf

template <typename T>
struct S2Args {
    SubPdfArgs<T> a;
    FvmLayerC layers[JF_MAX_LAYERS];
    SplineC<T> splines[kS2MaxSplines];
};

// S2 sub-pdf: chain of "f" layers; the first layer carries the plane<->sphere base chart.
template <typename T, int DIR>
__global__ void __launch_bounds__(256) s2_chain_kernel(const __grid_constant__ S2Args<T> g) {
    const SubPdfArgs<T>& a = g.a;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.B) return;
    T c0 = a.in[row * a.ld_in + 0], c1 = a.in[row * a.ld_in + 1];
    T logdet = a.logdet_in ? a.logdet_in[row] : T(0);
    const T* prow = a.params + row * a.sr;
    T zsq;
    int oor = 0, evals = 0, unconv = 0;
    if (DIR == JF_DIR_LOGPDF) {
        if (a.emb_out) {
            T e[3], dummy = 0;
            s2_to_embedding(c0, c1, e, dummy);
            a.emb_out[row * a.ld_emb + 0] = e[0];
            a.emb_out[row * a.ld_emb + 1] = e[1];
            a.emb_out[row * a.ld_emb + 2] = e[2];
        }
        for (int l = a.n_layers - 1; l >= 0; --l) {
            if (g.layers[l].kind == JF_LAYER_EXPMAP) v_layer<T>(true, c0, c1, logdet, g.layers[l], prow, a.sj, evals, unconv);
            else fvm_logpdf<T>(c0, c1, logdet, g.layers[l], g.splines, prow, a.sj, oor);
        }
        zsq = c0 * c0 + c1 * c1;
    } else {
        zsq = c0 * c0 + c1 * c1;
        for (int l = 0; l < a.n_layers; ++l) {
            if (g.layers[l].kind == JF_LAYER_EXPMAP) v_layer<T>(false, c0, c1, logdet, g.layers[l], prow, a.sj, evals, unconv);
            else fvm_sample<T>(c0, c1, logdet, g.layers[l], g.splines, prow, a.sj, oor);
        }
        if (a.emb_out) {
            T e[3], dummy = 0;
            s2_to_embedding(c0, c1, e, dummy);
            a.emb_out[row * a.ld_emb + 0] = e[0];
            a.emb_out[row * a.ld_emb + 1] = e[1];
            a.emb_out[row * a.ld_emb + 2] = e[2];
        }
    }
    a.out[row * a.ld_out + 0] = c0;
    a.out[row * a.ld_out + 1] = c1;
    if (!finite_(c0) || !finite_(c1) || !finite_(logdet)) status_add(a.status, JF_STATUS_NONFINITE, 1);
    if (oor) status_add(a.status, JF_STATUS_OUT_OF_RANGE, oor);
    if (unconv) status_add(a.status, JF_STATUS_UNCONVERGED, unconv);
    status_add_warp(a.status, JF_STATUS_ITERATIONS, evals);
    if (a.logdet_out) a.logdet_out[row] = logdet;
    if (a.logbase_out) {
        const T prev = a.logbase_in ? a.logbase_in[row] : T(0);
        a.logbase_out[row] = prev - T(0.5) * zsq - T(2) * T(kLogSqrt2Pi);
    }
}

}  // namespace jf
