// Micro-benchmark: peak FP64 / FP32 FMA issue rate of the CUDA cores (the second roof of DESIGN.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_peak tools/fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void __launch_bounds__(256) fma_kernel(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == (T)123456789) out[0] = s;   // never true; keeps the loop alive
}

template <typename T, int ILP>
double run(const char* name, int sms, int blocks_per_sm, int iters) {
    T* d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * blocks_per_sm;
    fma_kernel<T, ILP><<<grid, 256>>>(d, iters / 10, (T)1.0000001, (T)1e-9);
    cudaDeviceSynchronize();
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<T, ILP><<<grid, 256>>>(d, iters, (T)1.0000001, (T)1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)grid * 256 * ILP * iters;
        const double rate = fma / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    printf("{\"pipe\": \"%s\", \"tfma_per_s\": %.3f, \"tflops\": %.3f, \"blocks_per_sm\": %d, \"ilp\": %d}\n", name, best / 1e12,
           2 * best / 1e12, blocks_per_sm, ILP);
    cudaFree(d);
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d}\n", p.name, p.multiProcessorCount, clk / 1000);
    const int sms = p.multiProcessorCount;
    for (int bps : {2, 4, 8}) {
        double r64 = run<double, 8>("fp64", sms, bps, 20000);
        double r32 = run<float, 8>("fp32", sms, bps, 40000);
        printf("{\"fp64_fma_per_clk_per_sm_at_max_clock\": %.2f, \"fp32_fma_per_clk_per_sm_at_max_clock\": %.2f}\n",
               r64 / sms / (clk * 1e3), r32 / sms / (clk * 1e3));
    }
    run<double, 16>("fp64", sms, 4, 20000);
    run<float, 16>("fp32", sms, 4, 40000);
    // low occupancy: 256 threads per SM = 2 warps per scheduler (the marching kernels' regime)
    run<double, 8>("fp64_2warps_per_sched", sms, 1, 20000);
    run<double, 24>("fp64_2warps_per_sched", sms, 1, 20000);
    run<float, 8>("fp32_2warps_per_sched", sms, 1, 40000);
    run<float, 24>("fp32_2warps_per_sched", sms, 1, 40000);
    return 0;
}
