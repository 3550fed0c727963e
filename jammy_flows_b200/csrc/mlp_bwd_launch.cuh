// Host-side interface of the tensor-core generator backward (defined in mlp_bwd_inst.cu, used by api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jf {
struct BwArgs {
    const float* G; int64_t ldg;            // [P, B] param-major: element (p, row) at G[p * ldg + row]
    int P; int64_t B;
    const float* x; int64_t ldx; int in;    // generator input [B, in]
    const float* W1; const float* b1;       // [128, in], [128]
    const float* W2;                        // [P, 128]
    float* w2_tiles;                        // workspace: ceil(P/32) tiles [128 n][32 k]
    float* h_tiles;                         // workspace: ceil(B/32) tiles [144 n][32 k]
    float* dpre;                            // workspace: [B, 128]
    const float* row_scale;                 // optional [B]: the upstream gradient is G[p, row] * row_scale[row]
    float* fac;                             // workspace: [B, 128] (1 - h^2) * row_scale
    float* dW1; float* db1; float* dW2; float* db2;   // outputs (zeroed by the caller, accumulated)
    float* dx; int64_t lddx;                // optional [B, in]
    int n_splits;                           // bw_dw2: row ranges per parameter tile
};
int64_t mlp_bwd_workspace_bytes(int P, int64_t B);
// launches the five kernels (5 launches); the gradient outputs must be zeroed by the caller
int launch_mlp_bwd(BwArgs a, void* workspace, cudaStream_t st);
constexpr int kMlpBwdLaunches = 5;
}  // namespace jf
