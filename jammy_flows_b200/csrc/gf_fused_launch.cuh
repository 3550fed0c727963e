// Host-side interface of the fused generator + "g"-chain kernel (defined in gf_fused_inst.cu, used by api.cu).
#pragma once
#include "common.cuh"
#include "mlp_kernels.cuh"

namespace jf {

// int8 slices of the fused kernel's contraction (csrc/mlp_i8.cuh; error of the generated parameters relative to the
// scale of their W2 row: 7 slices 6e-15, 6 slices 3e-13, tools/ozaki_probe.py).  log_pdf: 6 -- the parameter error enters
// log p with O(1..30) sensitivity, 1e-11 against the 1e-10 contract, and the tensor pipe is co-critical with the FP64
// pipe there.  Sampling: 7 -- the root x(z) amplifies a parameter error by 1/pdf (2.6e-10 measured with 6 slices on the
// golden emb_e2s2e2_cond), and its tensor work hides behind the root finder anyway.
#ifndef JF_FUSED_NS_LOGPDF
#define JF_FUSED_NS_LOGPDF 6
#endif
#ifndef JF_FUSED_NS_SAMPLE
#define JF_FUSED_NS_SAMPLE 7
#endif
constexpr int kFuNSLogpdf = JF_FUSED_NS_LOGPDF, kFuNSSample = JF_FUSED_NS_SAMPLE;
// MMA tile width (N): tensor memory holds NS accumulators of TN columns plus the NS A slices of 32 columns (<= 512)
#ifndef JF_FUSED_TN_LOGPDF
#define JF_FUSED_TN_LOGPDF 48
#endif
#ifndef JF_FUSED_TN_SAMPLE
#define JF_FUSED_TN_SAMPLE 32
#endif
constexpr int kFuTNLogpdf = JF_FUSED_TN_LOGPDF, kFuTNSample = JF_FUSED_TN_SAMPLE;
constexpr int kFuK = 10;               // num_kde the column layout is built for: 3 K = 30 of the 36 slots of a (layer, dimension)
constexpr int kFuMaxD = 4;             // dimensions = worker column groups
constexpr int kFuMaxHH = 4;            // Householder reflections per layer

struct FuLayerC {
    int inv_type, has_offset, hh_iter, raw_off;
    double w_min, inv_w_max, n_min, n_max;
};

struct FuArgs {
    MlpArgs<double> m;              // input gather, W1, b1 (m.out unused)
    int n_layers, d;
    FuLayerC layers[JF_MAX_LAYERS]; // flow order
    const double* in;  int64_t ld_in;
    double* out;       int64_t ld_out;
    const double* logdet_in;  double* logdet_out;
    const double* logbase_in; double* logbase_out;
    int64_t* status;
    const unsigned char* wsB;       // [tiles][NS slices][TN x 128 B] in consumption order
    const double2* consts;          // [tiles * TN] (scale, b2) per fused column
};

int64_t fused_prep_bytes(int n_layers);
// slices W2 / b2 into the workspace in the consumption order of `direction`, and points a.wsB / a.consts at it
int launch_fused_prep(FuArgs& a, const double* W2, const double* b2, int direction, void* ws, bool run, cudaStream_t st);
int launch_fused(const FuArgs& a, int direction, cudaStream_t st);
// does the kernel of this direction fit into `smem_max` bytes of shared memory with `kin` generator inputs?
bool fused_fits(int direction, int kin, int smem_max);

}  // namespace jf
