// Downstream analysis step (reference example_analysis_script.ipynb cells 4-6): reliability
// percentile by radix select, then one fused elementwise kernel for mask -> NaN -> physical units -> magnitude / angles.
// Internal header.
#pragma once
#include "common.cuh"

namespace of3d {

// order-preserving map of IEEE bits to unsigned keys
__device__ __forceinline__ uint32_t sort_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint64_t sort_key(double v) {
    const uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

constexpr int kRadixBits = 11, kRadixBins = 1 << kRadixBits;

// One pass of an MSB-first radix select: histogram of the (up to kRadixBits) `bin_mask` bits at `shift` over the elements whose key
// matches `prefix` on the bits above them (`prefix_mask`).  Also counts NaNs (pass 0 only needs it).
template <typename T, typename Key>
__global__ void __launch_bounds__(256) radix_hist(const T* __restrict__ x, int64_t n, Key prefix, Key prefix_mask, int shift, unsigned bin_mask,
                                                   unsigned long long* __restrict__ hist, unsigned long long* __restrict__ n_nan) {
    __shared__ unsigned int sh[kRadixBins];
    for (int i = threadIdx.x; i < kRadixBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned int nan_local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T v = x[i];
        if (v != v) { ++nan_local; continue; }
        const Key k = sort_key(v);
        if ((k & prefix_mask) == prefix) atomicAdd(&sh[(unsigned)(k >> shift) & bin_mask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kRadixBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
    if (nan_local) atomicAdd(n_nan, (unsigned long long)nan_local);
}

// example_analysis_script.ipynb cell 5-6, per component:  m = v * mask;  m[m == 0] = nan;  m = m * scale / tscale
template <typename T>
__device__ __forceinline__ T mask_scale(T v, bool keep, T scale, T tscale) {
    T m = v * (keep ? T(1) : T(0));
    if (m == T(0)) m = nan("");
    return (m * scale) / tscale;
}

template <typename T, typename TR>
__global__ void __launch_bounds__(256) mask_derive(const T* __restrict__ vx, const T* __restrict__ vy, const T* __restrict__ vz,
                                                    const TR* __restrict__ rel, int64_t n, TR thresh, T sxy, T sz, T tscale,
                                                    T* __restrict__ ox, T* __restrict__ oy, T* __restrict__ oz,
                                                    T* __restrict__ mag, T* __restrict__ theta, T* __restrict__ phi) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool keep = rel[i] > thresh;                       // relMask = rel > relThresh (NaN -> False)
        const T x = mask_scale<T>(vx[i], keep, sxy, tscale);
        const T y = mask_scale<T>(vy[i], keep, sxy, tscale);
        ox[i] = x; oy[i] = y;
        const T xy2 = x * x + y * y;
        if (vz) {
            const T z = mask_scale<T>(vz[i], keep, sz, tscale);
            oz[i] = z;
            mag[i] = sqrt(xy2 + z * z);                          // sqrt(vx^2 + vy^2 + vz^2)
            phi[i] = atan(z / sqrt(xy2));                        // angle with respect to the xy plane
        } else {
            mag[i] = sqrt(xy2);
        }
        theta[i] = atan2(y, x);                                  // angle in the xy plane, -pi..pi
    }
}

}  // namespace of3d
