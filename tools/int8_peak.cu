// Measured tcgen05 kind::i8 throughput on this B200 (the roofline denominator of the int8-sliced generator kernels):
// one CTA per SM issues back-to-back M = 128 MMAs on resident operands (no loads), for several N and for the A operand
// in shared memory (SS) or tensor memory (TS).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/int8_peak tools/int8_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    for (int i = threadIdx.x; i < (4096 + N * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint64_t ad = desc(sb, 16 * 128, 128), bd = desc(sb + 4096, (N / 8) * 128, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (TS)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %3, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %4, p;\n}"
                             ::"r"(tmem), "r"(tmem + 256), "l"(bd), "r"(it), "r"(idesc) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %3, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %4, p;\n}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(it), "r"(idesc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar_a) : "memory");
        if (blockIdx.x == 0) cyc[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, bool TS>
void run(int sms, long long* cyc) {
    const int iters = 1 << 15;
    const int smem = 4096 + N * 32 + 1024;
    cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<N, TS><<<sms, 128, smem>>>(iters, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<N, TS><<<sms, 128, smem>>>(iters, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = 2.0 * 128 * N * 32 * iters * sms;
    printf("{\"kind\": \"i8\", \"M\": 128, \"N\": %d, \"A\": \"%s\", \"cycles_per_mma\": %.1f, \"tops\": %.1f, \"err\": \"%s\"}\n", N, TS ? "tmem" : "smem",
           (double)c / iters, ops / (ms * 1e-3) * 1e-12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* cyc;
    cudaMalloc(&cyc, 8);
    run<256, false>(sms, cyc); run<128, false>(sms, cyc); run<64, false>(sms, cyc); run<48, false>(sms, cyc); run<32, false>(sms, cyc);
    run<256, true>(sms, cyc); run<64, true>(sms, cyc); run<48, true>(sms, cyc); run<32, true>(sms, cyc);
    return 0;
}
