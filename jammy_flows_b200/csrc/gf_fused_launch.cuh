// Host-side interface of the fused generator + "g"-chain kernel (defined in gf_fused_inst.cu, used by api.cu).
#pragma once
#include "common.cuh"
#include "mlp_kernels.cuh"

namespace jf {

#ifndef JF_FUSED_NS
#define JF_FUSED_NS 7
#endif
constexpr int kFuNS = JF_FUSED_NS;     // int8 slices of the fused kernel's contraction (see csrc/mlp_i8.cuh)
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
    const unsigned char* wsB;       // [3 L tiles][NS slices][48 x 128 B] in consumption order
    const double2* consts;          // [3 L * 48] (scale, b2) per fused column
};

int64_t fused_prep_bytes(int n_layers);
// slices W2 / b2 into the workspace in the consumption order of `direction`, and points a.wsB / a.consts at it
int launch_fused_prep(FuArgs& a, const double* W2, const double* b2, int direction, void* ws, bool run, cudaStream_t st);
int launch_fused(const FuArgs& a, int direction, cudaStream_t st);

}  // namespace jf
