// FP64 pipe occupancy study (B200): DFMA throughput of ONE SM as a function of resident warps and of the number of
// independent chains per thread (ILP), plus the mix exp-like chain the layer kernels run.  One block on one SM, clock64.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* cyc) {
    const int iters = 4096;
    k<ILP><<<1, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    k<ILP><<<1, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / iters;                       // cycles per iteration (ILP DFMAs per thread)
    const double warp_instr_per_clk = (double)warps * ILP / per;   // per SM
    printf("warps %2d ILP %d: %.2f cycles/iter  -> %.3f warp-DFMA/clk/SM (peak 2.0), %.1f cycles per dependent DFMA\n", warps, ILP, per,
           warp_instr_per_clk, per / 1.0);
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    for (int warps : {4, 8, 16, 20, 24, 32}) {
        run<1>(warps, out, cyc);
        run<2>(warps, out, cyc);
        run<4>(warps, out, cyc);
        run<8>(warps, out, cyc);
    }
    return 0;
}
