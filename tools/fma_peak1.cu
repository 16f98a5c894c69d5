// single-warp-per-scheduler variant: 128 threads per SM
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void __launch_bounds__(128) k(T* out, int iters, T a, T b) {
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
    if (s == (T)123456789) out[0] = s;
}
template <typename T, int ILP> void run(const char* n, int grid, int iters) {
    T* d; cudaMalloc(&d, 64); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<T, ILP><<<grid, 128>>>(d, iters / 10, (T)1.0000001, (T)1e-9); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<T, ILP><<<grid, 128>>>(d, iters, (T)1.0000001, (T)1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"pipe\": \"%s\", \"warps_per_sched\": 1, \"ilp\": %d, \"tfma_per_s\": %.3f}\n", n, ILP, (double)grid * 128 * ILP * iters / (ms * 1e-3) / 1e12);
}
int main() { run<double, 8>("fp64", 148, 20000); run<double, 24>("fp64", 148, 20000); run<float, 8>("fp32", 148, 40000); run<float, 24>("fp32", 148, 40000); return 0; }
