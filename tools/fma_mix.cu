// Micro-benchmark: how much of the FP64 FMA rate survives when other instructions share the issue port.
// Each iteration issues 24 independent DFMAs (register operands) plus NI integer adds and NL shared-memory loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_mix tools/fma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int NL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) mix_kernel(double* out, int iters, double a, double b, int seed) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 1e-9 * i;
    __syncthreads();
    double acc[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) acc[i] = (double)(threadIdx.x + i);
    int q[4] = {seed, seed + 1, seed + 2, seed + 3};
    double l = 0;
    const double* sp = sh + (threadIdx.x & 31);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(a), "d"(b));
            if (i < NI) asm volatile("add.s32 %0, %0, %1;" : "+r"(q[i & 3]) : "r"(q[(i + 1) & 3]));
            if (i + 24 < NI) asm volatile("add.s32 %0, %0, %1;" : "+r"(q[(i + 2) & 3]) : "r"(q[(i + 3) & 3]));
            if (i < NL) { double t; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "l"(sp + 32 * i + (q[0] & 0)) ); l += t * 0; }
            if (i < -NL) { double t, t2; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t), "=d"(t2) : "l"(sh + 2 * (threadIdx.x & 31) + 64 * i + (q[0] & 0)) ); l += (t + t2) * 0; }
        }
    }
    double s = l + q[0] + q[1] + q[2] + q[3];
#pragma unroll
    for (int i = 0; i < 24; ++i) s += acc[i];
    if (s == 123456789.0) out[0] = s;
}

template <int NI, int NL, int WARPS>
void run(int sms, int iters) {
    double* d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix_kernel<NI, NL, WARPS><<<sms, WARPS * 32>>>(d, iters / 10, 1.0000001, 1e-9, 1);
    cudaDeviceSynchronize();
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        mix_kernel<NI, NL, WARPS><<<sms, WARPS * 32>>>(d, iters, 1.0000001, 1e-9, 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)sms * WARPS * 32 * 24 * iters / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    printf("{\"warps_per_sm\": %d, \"int_per_24_dfma\": %d, \"lds_per_24_dfma\": %d, \"tfma_per_s\": %.3f}\n", WARPS, NI, NL, best / 1e12);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, it = 20000;
    run<0, 0, 16>(sms, it); run<8, 0, 16>(sms, it); run<16, 0, 16>(sms, it); run<24, 0, 16>(sms, it); run<32, 0, 16>(sms, it); run<48, 0, 16>(sms, it);
    run<0, 2, 16>(sms, it); run<0, 4, 16>(sms, it); run<16, 2, 16>(sms, it);
    run<0, 0, 20>(sms, it); run<16, 2, 20>(sms, it); run<24, 2, 20>(sms, it);
    run<0, 0, 8>(sms, it); run<16, 2, 8>(sms, it);
    // negative NL: that many 128-bit shared loads (two doubles per lane) instead of 64-bit ones
    run<0, -1, 16>(sms, it); run<0, -2, 16>(sms, it); run<0, -4, 16>(sms, it); run<16, -2, 16>(sms, it);
    return 0;
}
