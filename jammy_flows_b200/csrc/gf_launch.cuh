// Launch dispatch of the Euclidean "g"-chain kernels.  Each (dtype, direction) pair is compiled in its own
// translation unit (gf_inst_*.cu) so that the library builds in parallel; api.cu only sees the declarations.
#pragma once
#include "subpdf_kernels.cuh"

namespace jf {

template <typename T, int D_, int DIR>
static int launch_gf_one(const GfChainArgs<T>& g, size_t smem_table, int kmax, cudaStream_t st) {
    // shared parameters: CTA-wide table; per-row parameters: 3*Kmax slots per thread -> pick the largest block that fits
    int threads = 256;
    size_t smem = smem_table;
    if (g.a.sr != 0) {
        while (threads > 32 && (size_t)3 * kmax * threads * sizeof(T) > 72 * 1024) threads >>= 1;
        smem = (size_t)3 * kmax * threads * sizeof(T);
    }
    const int64_t blocks = (g.a.B + threads - 1) / threads;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gf_chain_kernel<T, D_, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    gf_chain_kernel<T, D_, DIR><<<(unsigned)blocks, threads, smem, st>>>(g);
    return JF_OK;
}

// returns JF_OK after enqueueing (the caller checks cudaGetLastError)
template <typename T, int DIR>
int launch_gf_dir(const GfChainArgs<T>& g, int d, int kmax, size_t smem_table, cudaStream_t st);

#define JF_GF_LAUNCH_DIR_BODY(T, DIR)                                                                         \
    template <> int launch_gf_dir<T, DIR>(const GfChainArgs<T>& g, int d, int kmax, size_t smem, cudaStream_t st) { \
        switch (d) {                                                                                          \
            case 1: return launch_gf_one<T, 1, DIR>(g, smem, kmax, st);                                       \
            case 2: return launch_gf_one<T, 2, DIR>(g, smem, kmax, st);                                       \
            case 3: return launch_gf_one<T, 3, DIR>(g, smem, kmax, st);                                       \
            case 4: return launch_gf_one<T, 4, DIR>(g, smem, kmax, st);                                       \
            case 5: return launch_gf_one<T, 5, DIR>(g, smem, kmax, st);                                       \
            case 6: return launch_gf_one<T, 6, DIR>(g, smem, kmax, st);                                       \
            case 8: return launch_gf_one<T, 8, DIR>(g, smem, kmax, st);                                       \
            case 10: return launch_gf_one<T, 10, DIR>(g, smem, kmax, st);                                     \
            default: break;                                                                                   \
        }                                                                                                     \
        return launch_gf_one<T, 0, DIR>(g, smem, kmax, st);                                                   \
    }

}  // namespace jf
