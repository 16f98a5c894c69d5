// Micro-benchmark: FP64 tensor-core (DMMA, mma.sync f64) issue rate on B200, alone and next to DFMA work.
// Question it answers: is the DMMA pipe a second FP64 roof beside the CUDA-core DFMA pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                 "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
                 "{%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// MODE 0: m8n8k4 only; 1: m16n8k16 only; 2: DFMA only; 3: m8n8k4 + DFMA in the same warp (NF DFMA per DMMA);
// 4: m16n8k8 only
template <int MODE, int ILP, int NF>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double c[ILP][4];
    double f[ILP * (NF > 0 ? NF : 1)];
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = a + i * 1e-9 + threadIdx.x * 1e-12;
#pragma unroll
    for (int i = 0; i < 4; ++i) bv[i] = b + i * 1e-9;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 2 * i; c[i][3] = 3 * i;
    }
#pragma unroll
    for (int i = 0; i < ILP * (NF > 0 ? NF : 1); ++i) f[i] = threadIdx.x * 0.5 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0 || MODE == 3) dmma884(c[i][0], c[i][1], av[0], bv[0]);
            if (MODE == 1) dmma16816(c[i], av, bv);
            if (MODE == 4) { double a4[4] = {av[0], av[1], av[2], av[3]}; double b2[2] = {bv[0], bv[1]}; dmma1688(c[i], a4, b2); }
            if (MODE == 2 || MODE == 3) {
#pragma unroll
                for (int j = 0; j < NF; ++j)
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i * NF + j]) : "d"(a), "d"(b));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
#pragma unroll
    for (int i = 0; i < ILP * (NF > 0 ? NF : 1); ++i) s += f[i];
    if (s == 123456789.0) out[0] = s;
}

template <int MODE, int ILP, int NF>
void run(const char* name, int sms, int bps, int iters) {
    double* d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * bps;
    k<MODE, ILP, NF><<<grid, 256>>>(d, iters / 10, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    double best_ms = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE, ILP, NF><<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best_ms) best_ms = ms;
    }
    const double warps = (double)grid * 8;
    const double per_mma = MODE == 0 || MODE == 3 ? 256.0 : MODE == 1 ? 2048.0 : MODE == 4 ? 1024.0 : 0.0;
    const double mma_fma = warps * ILP * iters * per_mma;
    const double dfma = (MODE == 2 || MODE == 3) ? warps * 32 * ILP * NF * iters : 0.0;
    printf("{\"case\": \"%s\", \"blocks_per_sm\": %d, \"ilp\": %d, \"dfma_per_mma\": %d, \"ms\": %.3f, "
           "\"dmma_tfma_per_s\": %.3f, \"dfma_tfma_per_s\": %.3f, \"sum_tfma_per_s\": %.3f}\n",
           name, bps, ILP, NF, best_ms, mma_fma / best_ms / 1e9, dfma / best_ms / 1e9, (mma_fma + dfma) / best_ms / 1e9);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d}\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<2, 8, 1>("dfma_only", sms, 4, 20000);
    run<0, 8, 0>("dmma_m8n8k4", sms, 4, 20000);
    run<0, 8, 0>("dmma_m8n8k4", sms, 2, 20000);
    run<0, 4, 0>("dmma_m8n8k4", sms, 1, 20000);
    run<4, 8, 0>("dmma_m16n8k8", sms, 4, 5000);
    run<1, 8, 0>("dmma_m16n8k16", sms, 4, 2500);
    run<1, 4, 0>("dmma_m16n8k16", sms, 2, 2500);
    run<3, 8, 1>("dmma884+dfma", sms, 4, 10000);
    run<3, 8, 2>("dmma884+dfma", sms, 4, 10000);
    run<3, 8, 4>("dmma884+dfma", sms, 4, 5000);
    run<3, 4, 8>("dmma884+dfma", sms, 4, 5000);
    return 0;
}
