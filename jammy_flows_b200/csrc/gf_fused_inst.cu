// instantiation + launcher of the fused generator + "g"-chain kernel (csrc/gf_fused.cuh); its own translation unit so
// that the library builds in parallel
#include "gf_fused_launch.cuh"
#include "gf_fused.cuh"

namespace jf {

template <int NS, int TN, int DIR, int KR>
static int launch_fused_one(const FuArgs& a, int smem_max, int sms, cudaStream_t st) {
    using G = FuCfg<NS, TN, DIR>;
    const int Kin = a.m.dims[0];
    int n_stages = 3;
    while (n_stages > 1 && !G::fits(Kin, n_stages, smem_max)) --n_stages;
    if (!G::fits(Kin, n_stages, smem_max)) return JF_ERR_UNSUPPORTED;
    const int smem = G::smem_bytes(Kin, n_stages);
    cudaError_t e = cudaFuncSetAttribute(gf_fused_kernel<NS, TN, DIR, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t blocks = (a.m.B + kI8Rows - 1) / kI8Rows;
    const unsigned grid = (unsigned)(blocks < sms ? blocks : sms);
    gf_fused_kernel<NS, TN, DIR, KR><<<grid, kFuThreads, smem, st>>>(a, n_stages);
    return JF_OK;
}

int64_t fused_prep_bytes(int n_layers) {
    const int64_t a = fu_prep_bytes<kFuNSLogpdf, kFuTNLogpdf>(n_layers), b = fu_prep_bytes<kFuNSSample, kFuTNSample>(n_layers);
    return a > b ? a : b;
}

int launch_fused_prep(FuArgs& a, const double* W2, const double* b2, int direction, void* ws, bool run, cudaStream_t st) {
    const bool lp = direction == JF_DIR_LOGPDF;
    const int n_tiles = (lp ? FuCfg<kFuNSLogpdf, kFuTNLogpdf>::kTPL : FuCfg<kFuNSSample, kFuTNSample>::kTPL) * a.n_layers;
    const int ns = lp ? kFuNSLogpdf : kFuNSSample, tn = lp ? kFuTNLogpdf : kFuTNSample;
    a.wsB = (const unsigned char*)ws;
    a.consts = reinterpret_cast<const double2*>((const unsigned char*)ws + (size_t)n_tiles * ns * tn * kI8H);
    if (run) {
        if (lp) fu_prep_kernel<kFuNSLogpdf, kFuTNLogpdf><<<n_tiles, 128, 0, st>>>(a, W2, b2, direction, (unsigned char*)ws);
        else fu_prep_kernel<kFuNSSample, kFuTNSample><<<n_tiles, 128, 0, st>>>(a, W2, b2, direction, (unsigned char*)ws);
    }
    return JF_OK;
}

bool fused_fits(int direction, int kin, int smem_max) {
    return direction == JF_DIR_LOGPDF ? FuCfg<kFuNSLogpdf, kFuTNLogpdf, JF_DIR_LOGPDF>::fits(kin, 1, smem_max)
                                      : FuCfg<kFuNSSample, kFuTNSample, JF_DIR_SAMPLE>::fits(kin, 1, smem_max);
}

int launch_fused(const FuArgs& a, int direction, cudaStream_t st) {
    int dev = 0, sms = 0, smem_max = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return (int)e;
    const bool k8 = a.m.dims[0] <= 8;
    if (direction == JF_DIR_LOGPDF)
        return k8 ? launch_fused_one<kFuNSLogpdf, kFuTNLogpdf, JF_DIR_LOGPDF, 8>(a, smem_max, sms, st)
                  : launch_fused_one<kFuNSLogpdf, kFuTNLogpdf, JF_DIR_LOGPDF, 16>(a, smem_max, sms, st);
    return k8 ? launch_fused_one<kFuNSSample, kFuTNSample, JF_DIR_SAMPLE, 8>(a, smem_max, sms, st)
              : launch_fused_one<kFuNSSample, kFuTNSample, JF_DIR_SAMPLE, 16>(a, smem_max, sms, st);
}

}  // namespace jf
