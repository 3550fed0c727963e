// instantiation unit: general "g"-chain kernels (non-default layer options), both dtypes and directions
#include "gfx.cuh"
namespace jf {

template <typename T>
int launch_gfx(const GfxChainArgs<T>& g, int direction, int kmax, cudaStream_t st) {
    int threads = 256;
    while (threads > 32 && (size_t)kGfxFields * kmax * threads * sizeof(T) > 64 * 1024) threads >>= 1;
    const size_t smem = (size_t)kGfxFields * kmax * threads * sizeof(T);
    if (smem > 200 * 1024) return JF_ERR_UNSUPPORTED;
    const int64_t blocks = (g.a.B + threads - 1) / threads;
    if (blocks == 0) return JF_OK;
    cudaError_t e;
    if (direction == JF_DIR_LOGPDF) {
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(gfx_chain_kernel<T, JF_DIR_LOGPDF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        gfx_chain_kernel<T, JF_DIR_LOGPDF><<<(unsigned)blocks, threads, smem, st>>>(g);
    } else {
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(gfx_chain_kernel<T, JF_DIR_SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        gfx_chain_kernel<T, JF_DIR_SAMPLE><<<(unsigned)blocks, threads, smem, st>>>(g);
    }
    return JF_OK;
}

template int launch_gfx<double>(const GfxChainArgs<double>&, int, int, cudaStream_t);
template int launch_gfx<float>(const GfxChainArgs<float>&, int, int, cudaStream_t);

}  // namespace jf
