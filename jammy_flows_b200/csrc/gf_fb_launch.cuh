// Host-side interface of the forward + backward "g"-chain kernel (csrc/gf_fb.cuh; instantiated in gf_fb_inst.cu, used by api.cu).
#pragma once
#include "gf.cuh"
#include "subpdf_args.cuh"

namespace jf {

template <typename T>
struct GfFbArgs {
    SubPdfArgs<T> a;           // in = x, params = raw per-row parameters, out = base z (may be NULL), logdet_out / logbase_out (may be NULL)
    const T* grad_logp;        // [B] upstream gradient of log_pdf (NULL = 1)
    T* grad_params;            // indexed like params
    T* grad_x; int64_t ld_gx;  // optional [B, d]: d log_pdf / d x (times grad_logp); sampling backward: cotangent of the base z
    const T* grad_out_x; int64_t ld_go;   // sampling backward only: cotangent of the sample x [B, d] (NULL = 0)
    int kmax, hh_max;          // slot / exchange geometry
    GfLayerC<T> layers[JF_MAX_LAYERS];
};

// row groups (of 32 rows) per block: about 256-320 threads
__host__ __device__ constexpr int fb_groups(int D) { return D >= 8 ? 1 : (D >= 4 ? 2 : (D >= 2 ? 4 : 8)); }
__host__ __device__ constexpr int fb_threads(int D) { return 32 * D * fb_groups(D); }
#ifndef JF_FB_REGS32
#define JF_FB_REGS32 96
#endif
// resident blocks per SM the register allocation aims at: 96 registers per thread in fp32, 128 in fp64 (the unrolled register-resident mixture)
__host__ __device__ constexpr int fb_min_blocks(int D, size_t elem) { return (int)(65536 / ((elem == 4 ? JF_FB_REGS32 : 128) * fb_threads(D))); }
template <typename T>
__host__ __device__ constexpr size_t fb_smem_bytes(int D, int kmax, int hh_max) {
    return ((size_t)3 * kmax * fb_threads(D) + (size_t)fb_groups(D) * (hh_max * D + 2 * D) * 32) * sizeof(T);
}

// launches gf_chain_fb_kernel<T, d, 0> (log_pdf forward + backward) / <T, d, 1> (backward of the sampling direction);
// return a JF_ERR_* / cudaError code
template <typename T> int launch_gf_fb(const GfFbArgs<T>& g, cudaStream_t st);
template <typename T> int launch_gf_sbwd(const GfFbArgs<T>& g, cudaStream_t st);
template <typename T> int launch_gf_fwd(const GfFbArgs<T>& g, cudaStream_t st);      // <T, d, 2>: log_pdf forward only

}  // namespace jf
