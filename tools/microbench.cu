// Microbenchmarks for the roofline denominators that MEASURED_PEAKS.json does not carry:
// FP64 DFMA peak, FP64 DMMA (mma.sync f64) peak, both together, and fp64/fp32 special-function throughput.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/microbench tools/microbench.cu
// Run on the B200 box only (prints one JSON object).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}

// DMMA m8n8k4: per warp 8*8*4 = 256 FMA
template <int ILP>
__global__ void k_dmma884(double* out, int iters) {
    double c0[ILP], c1[ILP];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

// DMMA m16n8k8 (sm_90+): per warp 16*8*8 = 1024 FMA; a: 4 regs, b: 2 regs, c: 4 regs
template <int ILP>
__global__ void k_dmma1688(double* out, int iters) {
    double c[ILP][4];
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-4, b1 = b0 * 0.5;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 3 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// DMMA + DFMA interleaved in the same warp: do they share a pipe?
template <int ILP>
__global__ void k_mix(double* out, int iters, double fa, double fb) {
    double c0[ILP], c1[ILP], acc[ILP];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; acc[i] = i * 0.5; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
            // 8 DFMA per thread = same FMA count as one m8n8k4 per warp (256/32)
            acc[i] = fma(acc[i], fa, fb); acc[i] = fma(acc[i], fa, fb); acc[i] = fma(acc[i], fa, fb); acc[i] = fma(acc[i], fa, fb);
            acc[i] = fma(acc[i], fa, fb); acc[i] = fma(acc[i], fa, fb); acc[i] = fma(acc[i], fa, fb); acc[i] = fma(acc[i], fa, fb);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i] + acc[i];
    if (s == 12345.678) out[0] = s;
}

enum { F_EXP = 0, F_LOG, F_DIV, F_ERFINV, F_SINCOS, F_ACOS, F_TANH, F_LOG1P, F_RCP, F_SQRT };

template <typename T, int F, int ILP>
__global__ void k_special(T* out, int iters, T seed) {
    T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + (T)(threadIdx.x & 31) * (T)1e-3 + (T)i * (T)1e-2;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            T x = v[i];
            if (F == F_EXP) x = exp(-x) + (T)0.5;                  // stays in (0.5,1.5)
            else if (F == F_LOG) x = log(x) + (T)1.5;                // x in ~(1,2)
            else if (F == F_DIV) x = (T)1.0 / x + (T)0.5;
            else if (F == F_RCP) x = (T)1.0 / x + (T)0.5;
            else if (F == F_ERFINV) x = erfinv(x * (T)0.3) + (T)0.7;
            else if (F == F_SINCOS) { T s, c; sincos(x, &s, &c); x = s * c + (T)1.0; }
            else if (F == F_ACOS) x = acos(x * (T)0.4) * (T)0.5 + (T)0.3;
            else if (F == F_TANH) x = tanh(x) + (T)0.5;
            else if (F == F_LOG1P) x = log1p(x) + (T)0.5;
            else if (F == F_SQRT) x = sqrt(x) + (T)0.5;
            v[i] = x;
        }
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == (T)12345.678) out[0] = s;
}

template <typename Launch>
static double time_ms(Launch&& l, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    l();  // warm up
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        l();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* dout; CK(cudaMalloc(&dout, 1024));
    float* fout = (float*)dout;
    const int threads = 512, blocks_per_sm = 4;
    const int grid = sms * blocks_per_sm;
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);

    {   // DFMA
        constexpr int ILP = 8;
        double ms = time_ms([&] { k_dfma<ILP><<<grid, threads>>>(dout, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * grid * threads * (double)iters * ILP;
        printf(", \"dfma_tflops\": %.3f", flops / ms * 1e-9);
    }
    {   // FFMA
        constexpr int ILP = 8;
        double ms = time_ms([&] { k_ffma<ILP><<<grid, threads>>>(fout, iters, 1.0000001f, 1e-9f); });
        double flops = 2.0 * grid * threads * (double)iters * ILP;
        printf(", \"ffma_tflops\": %.3f", flops / ms * 1e-9);
    }
    {   // DMMA m8n8k4
        constexpr int ILP = 8;
        double ms = time_ms([&] { k_dmma884<ILP><<<grid, threads>>>(dout, iters); });
        double flops = 2.0 * 256.0 * grid * (threads / 32) * (double)iters * ILP;
        printf(", \"dmma_m8n8k4_tflops\": %.3f", flops / ms * 1e-9);
    }
    {   // DMMA m16n8k8
        constexpr int ILP = 4;
        double ms = time_ms([&] { k_dmma1688<ILP><<<grid, threads>>>(dout, iters); });
        double flops = 2.0 * 1024.0 * grid * (threads / 32) * (double)iters * ILP;
        printf(", \"dmma_m16n8k8_tflops\": %.3f", flops / ms * 1e-9);
    }
    {   // mixed
        constexpr int ILP = 8;
        double ms = time_ms([&] { k_mix<ILP><<<grid, threads>>>(dout, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 2.0 * 256.0 * grid * (threads / 32) * (double)iters * ILP;
        printf(", \"dmma_plus_dfma_tflops\": %.3f", flops / ms * 1e-9);
    }
    const int it2 = 512;
#define SPECIAL(T, F, name)                                                                              \
    {                                                                                                    \
        constexpr int ILP = 4;                                                                           \
        double ms = time_ms([&] { k_special<T, F, ILP><<<grid, threads>>>((T*)dout, it2, (T)0.9); });    \
        double ops = (double)grid * threads * (double)it2 * ILP;                                         \
        printf(", \"" name "_gops\": %.2f", ops / ms * 1e-6);                                            \
    }
    SPECIAL(double, F_EXP, "f64_exp")
    SPECIAL(double, F_LOG, "f64_log")
    SPECIAL(double, F_DIV, "f64_div")
    SPECIAL(double, F_ERFINV, "f64_erfinv")
    SPECIAL(double, F_SINCOS, "f64_sincos")
    SPECIAL(double, F_ACOS, "f64_acos")
    SPECIAL(double, F_TANH, "f64_tanh")
    SPECIAL(double, F_LOG1P, "f64_log1p")
    SPECIAL(double, F_SQRT, "f64_sqrt")
    SPECIAL(float, F_EXP, "f32_exp")
    SPECIAL(float, F_LOG, "f32_log")
    SPECIAL(float, F_DIV, "f32_div")
    SPECIAL(float, F_ERFINV, "f32_erfinv")
    SPECIAL(float, F_TANH, "f32_tanh")
    printf("}\n");
    return 0;
}
