// launcher of the tensor-core backward of the parameter generator (csrc/mlp_bwd.cuh); its own translation unit so that
// the library builds in parallel
#include <cudaTypedefs.h>
#include "mlp_bwd.cuh"
#include "../../include/jammy_b200.h"

namespace jf {

static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

int64_t mlp_bwd_workspace_bytes(int P, int64_t B) {
    const int64_t w2 = (int64_t)((P + kBwKC - 1) / kBwKC) * kDhTile;
    const int64_t ht = ((B + kW2KC - 1) / kW2KC) * 2 * (int64_t)kW2TileB;      // an even number of 32-row h tiles
    return align256(w2) + align256(ht) + 2 * align256(B * kBwH * 4);
}

int launch_mlp_bwd(BwArgs a, void* workspace, cudaStream_t st) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    char* ws = (char*)workspace;
    const int n_ptiles32 = (a.P + kBwKC - 1) / kBwKC;
    const int64_t n_rtiles32 = (a.B + kW2KC - 1) / kW2KC * 2;     // even: bw_dw2 reads the h tiles in pairs (a tile past B is zero)
    a.w2_tiles = (float*)ws;
    a.h_tiles = (float*)(ws + align256((int64_t)n_ptiles32 * kDhTile));
    a.dpre = (float*)((char*)a.h_tiles + align256(n_rtiles32 * (int64_t)kW2TileB));
    a.fac = (float*)((char*)a.dpre + align256(a.B * kBwH * 4));
    bw_w2_tiles_kernel<<<n_ptiles32, 256, 0, st>>>(a.W2, a.P, a.w2_tiles);
    {
        const int smem = (a.in * kBwH + kBwKC * (a.in | 1)) * 4;
        e = cudaFuncSetAttribute(bw_h_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        const int64_t grid = n_rtiles32 < 4 * (int64_t)sms ? n_rtiles32 : 4 * (int64_t)sms;       // persistent over the tiles
        bw_h_tiles_kernel<<<(unsigned)grid, 256, smem, st>>>(a, n_rtiles32);
    }
    // tensor maps of G [P rows][B columns] fp32 in the 128-byte swizzle, zero fill outside: boxes of 32 columns x 32 rows
    // (bw_dh: MN-major A operand) and 32 columns x 128 rows (bw_dw2: K-major A operand)
    CUtensorMap tmapG32, tmapG128;
    {
        static PFN_cuTensorMapEncodeTiled encode = [] {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
                q != cudaDriverEntryPointSuccess)
                fn = nullptr;
            return (PFN_cuTensorMapEncodeTiled)fn;
        }();
        if (encode == nullptr) return JF_ERR_UNSUPPORTED;
        const cuuint64_t gdim[2] = {(cuuint64_t)a.B, (cuuint64_t)a.P};
        const cuuint64_t gstride[1] = {(cuuint64_t)a.ldg * 4};
        const cuuint32_t estr[2] = {1, 1};
        const cuuint32_t box32[2] = {32, 32}, box128[2] = {32, 128};
        CUresult r = encode(&tmapG32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.G, gdim, gstride, box32, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return JF_ERR_BAD_ARG;
        r = encode(&tmapG128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.G, gdim, gstride, box128, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return JF_ERR_BAD_ARG;
    }
    e = cudaFuncSetAttribute(bw_dh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDhSmem);
    if (e != cudaSuccess) return (int)e;
    bw_dh_kernel<<<(unsigned)((a.B + 127) / 128), kBwThreads, kDhSmem, st>>>(a, tmapG32);
    {
        // parameter tiles x row ranges: three waves of CTAs, at least 8 chunks of rows each
        const int n_pt = (a.P + 127) / 128;
        const int64_t n_chunks64 = n_rtiles32 / 2;
        int64_t splits = (3 * (int64_t)sms) / n_pt;                 // at most three full waves of CTAs (one CTA per SM)
        if (splits > (n_chunks64 + 7) / 8) splits = (n_chunks64 + 7) / 8;
        if (splits < 1) splits = 1;
        a.n_splits = (int)splits;
        e = cudaFuncSetAttribute(bw_dw2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW2Smem);
        if (e != cudaSuccess) return (int)e;
        bw_dw2_kernel<<<dim3((unsigned)n_pt, (unsigned)splits), kBwThreads, kW2Smem, st>>>(a, tmapG128);
    }
    {
        const int ldxs = ((a.in + 3) / 4) * 4 + 4 * ((kSmPer + 3) / 4);
        const int smem = (kSmRows * kSmLdD + kSmRows * ldxs + kBwH * (a.in | 1)) * 4;
        e = cudaFuncSetAttribute(bw_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        const int64_t n_tiles = (a.B + kSmRows - 1) / kSmRows;
        const int64_t grid = n_tiles < 2 * (int64_t)sms ? n_tiles : 2 * (int64_t)sms;      // persistent: two blocks per SM
        bw_small_kernel<<<(unsigned)grid, 256, smem, st>>>(a);
    }
    return 0;
}

}  // namespace jf
