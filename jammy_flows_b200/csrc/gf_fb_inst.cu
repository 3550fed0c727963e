// instantiations + launcher of the forward + backward "g"-chain kernel (csrc/gf_fb.cuh); its own translation unit so that
// the library builds in parallel
#include "gf_fb.cuh"

#ifndef JF_FB_MODE
#define JF_FB_MODE 0
#endif

namespace jf {

template <typename T, int D>
static int launch_fb_d(const GfFbArgs<T>& g, cudaStream_t st) {
    const size_t smem = fb_smem_bytes<T>(D, g.kmax, g.hh_max);
    if (smem > 200 * 1024) return JF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gf_chain_fb_kernel<T, D, JF_FB_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int rows_per_block = 32 * fb_groups(D);
    const int64_t blocks = (g.a.B + rows_per_block - 1) / rows_per_block;
    gf_chain_fb_kernel<T, D, JF_FB_MODE><<<(unsigned)blocks, fb_threads(D), smem, st>>>(g);
    return JF_OK;
}

#if JF_FB_MODE == 0
#define JF_FB_LAUNCH launch_gf_fb
#elif JF_FB_MODE == 1
#define JF_FB_LAUNCH launch_gf_sbwd
#else
#define JF_FB_LAUNCH launch_gf_fwd
#endif
template <typename T>
int JF_FB_LAUNCH(const GfFbArgs<T>& g, cudaStream_t st) {
    switch (g.a.d) {
#define JF_FB_CASE(D) case D: return launch_fb_d<T, D>(g, st);
        JF_FB_CASE(1) JF_FB_CASE(2) JF_FB_CASE(3) JF_FB_CASE(4) JF_FB_CASE(5) JF_FB_CASE(6) JF_FB_CASE(7) JF_FB_CASE(8)
        JF_FB_CASE(9) JF_FB_CASE(10) JF_FB_CASE(11) JF_FB_CASE(12) JF_FB_CASE(13) JF_FB_CASE(14) JF_FB_CASE(15) JF_FB_CASE(16)
#undef JF_FB_CASE
        default: return JF_ERR_UNSUPPORTED;
    }
}
template int JF_FB_LAUNCH<float>(const GfFbArgs<float>&, cudaStream_t);
template int JF_FB_LAUNCH<double>(const GfFbArgs<double>&, cudaStream_t);

}  // namespace jf
