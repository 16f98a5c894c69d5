// Generic (any tap count) kernels: one output element per thread, one pass per launch.
// They are the always-available fallback for tap counts without a specialised marching
// kernel, and -- with EXACT = true -- the bit-exact mode: scipy's paired summation order
// (ni_filters.c NI_Correlate1D, restated in oracle/lk_oracle.py:correlate1d_nearest) with
// individually rounded multiplies and adds.  Internal header.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "solve.cuh"

namespace of3d {

template <typename T> struct Rn;
template <> struct Rn<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
};
template <> struct Rn<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
};

// out[i] = sum_k w[k] * in[clamp(i + k - r)] along one axis of a contiguous volume.
// `len` = extent of the axis, `stride` = element stride of the axis.
template <typename T, bool EXACT>
__global__ void __launch_bounds__(256) corr_axis_generic(const T* __restrict__ in, T* __restrict__ out, int64_t n,
                                                         int64_t len, int64_t stride, const Filt<T> f) {
    const int r = f.n / 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pos = (i / stride) % len;
        const T* base = in + (i - pos * stride);
        auto at = [&](int64_t p) -> T {
            p = p < 0 ? 0 : (p >= len ? len - 1 : p);
            return __ldg(base + p * stride);
        };
        T acc;
        if (EXACT) {
            if (f.sym > 0) {
                acc = Rn<T>::mul(at(pos), f.w[r]);
                for (int l = -r; l < 0; ++l) acc = Rn<T>::add(acc, Rn<T>::mul(Rn<T>::add(at(pos + l), at(pos - l)), f.w[r + l]));
            } else if (f.sym < 0) {
                acc = Rn<T>::mul(at(pos), f.w[r]);
                for (int l = -r; l < 0; ++l) acc = Rn<T>::add(acc, Rn<T>::mul(Rn<T>::sub(at(pos + l), at(pos - l)), f.w[r + l]));
            } else {
                const int r2 = f.n - r - 1;
                acc = Rn<T>::mul(at(pos + r2), f.w[r + r2]);
                for (int l = -r; l < r2; ++l) acc = Rn<T>::add(acc, Rn<T>::mul(at(pos + l), f.w[r + l]));
            }
        } else {
            acc = T(0);
            for (int k = 0; k < f.n; ++k) acc = fma(f.w[k], at(pos + k - r), acc);
        }
        out[i] = acc;
    }
}

// Temporal derivative of the centre frame (calc_flow.py:276-278 / 113-115) and widening of the
// centre frame to the compute type (calc_flow.py:225 / 67).  frames.p[k] = frame c - r + k.
// paired (8/16-bit integer frames, antisymmetric taps with a zero centre): dt0 = sum_{l=1..r} w[r+l] (x[r+l] - x[r-l]) with
// the differences formed exactly in integers and accumulated by FMA in ascending l -- the arithmetic of the fused z march
// (kernels_tz.cuh), so that the two-stage and the fused pipelines agree bit for bit.
template <typename Tin, typename T, bool EXACT>
__global__ void __launch_bounds__(256) temporal_generic(const FramePtrs frames, const Filt<T> f, T* __restrict__ ic,
                                                        T* __restrict__ dt0, int64_t n, const int paired) {
    const int r = f.n / 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        auto at = [&](int k) -> T { return (T) __ldg(reinterpret_cast<const Tin*>(frames.p[k]) + i); };
        const T c = at(r);
        T acc;
        if (EXACT) {
            acc = Rn<T>::mul(c, f.w[r]);
            if (f.sym > 0)
                for (int l = -r; l < 0; ++l) acc = Rn<T>::add(acc, Rn<T>::mul(Rn<T>::add(at(r + l), at(r - l)), f.w[r + l]));
            else if (f.sym < 0)
                for (int l = -r; l < 0; ++l) acc = Rn<T>::add(acc, Rn<T>::mul(Rn<T>::sub(at(r + l), at(r - l)), f.w[r + l]));
            else {
                acc = Rn<T>::mul(at(f.n - 1), f.w[f.n - 1]);
                for (int k = 0; k < f.n - 1; ++k) acc = Rn<T>::add(acc, Rn<T>::mul(at(k), f.w[k]));
            }
        } else {
            acc = T(0);
            bool done = false;
            if constexpr (std::is_integral<Tin>::value && sizeof(Tin) <= 2) {
                if (paired) {
                    for (int l = 1; l <= r; ++l) {
                        const int d = (int)__ldg(reinterpret_cast<const Tin*>(frames.p[r + l]) + i) - (int)__ldg(reinterpret_cast<const Tin*>(frames.p[r - l]) + i);
                        acc = fma(f.w[r + l], (T)d, acc);
                    }
                    done = true;
                }
            }
            if (!done)
                for (int k = 0; k < f.n; ++k) acc = fma(f.w[k], at(k), acc);
        }
        ic[i] = c;
        dt0[i] = acc;
    }
}

// 16-byte stores of a register array (dst is 16-byte aligned)
template <int VEC>
__device__ __forceinline__ void store_vec(double* dst, const double (&v)[VEC]) {
#pragma unroll
    for (int j = 0; j < VEC; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(v[j], v[j + 1]);
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* dst, const float (&v)[VEC]) {
    if (VEC % 4 == 0) {
#pragma unroll
        for (int j = 0; j + 3 < VEC; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j + 1 < VEC; j += 2) *reinterpret_cast<float2*>(dst + j) = make_float2(v[j], v[j + 1]);
    }
}

// Vectorised variant for the marching pipeline: every thread converts VEC = 16 / sizeof(Tin) consecutive voxels
// (one 16-byte load per frame), so the kernel runs at HBM speed instead of being bound by 2-byte load instructions.
// Requires 16-byte aligned frame pointers; the scalar kernel above handles unaligned frames and the tail.
template <typename Tin, typename T>
__global__ void __launch_bounds__(256) temporal_vec(const FramePtrs frames, const Filt<T> f, T* __restrict__ ic,
                                                    T* __restrict__ dt0, int64_t nvec, const int paired) {
    constexpr int VEC = 16 / sizeof(Tin);
    static_assert(VEC % 2 == 0, "vector width");
    const int r = f.n / 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        T acc[VEC], c[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = T(0);
        if constexpr (std::is_integral<Tin>::value && sizeof(Tin) <= 2) {
            if (paired) {                                        // see temporal_generic
                const uint4 rc = __ldg(reinterpret_cast<const uint4*>(frames.p[r]) + i);
                const Tin* vc = reinterpret_cast<const Tin*>(&rc);
#pragma unroll
                for (int j = 0; j < VEC; ++j) c[j] = (T)vc[j];
                for (int l = 1; l <= r; ++l) {
                    const uint4 ra = __ldg(reinterpret_cast<const uint4*>(frames.p[r + l]) + i);
                    const uint4 rb = __ldg(reinterpret_cast<const uint4*>(frames.p[r - l]) + i);
                    const Tin* va = reinterpret_cast<const Tin*>(&ra);
                    const Tin* vb = reinterpret_cast<const Tin*>(&rb);
                    const T w = f.w[r + l];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[j] = fma(w, (T)((int)va[j] - (int)vb[j]), acc[j]);
                }
                store_vec<VEC>(ic + i * VEC, c);
                store_vec<VEC>(dt0 + i * VEC, acc);
                continue;
            }
        }
        for (int k = 0; k < f.n; ++k) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(frames.p[k]) + i);
            const Tin* v = reinterpret_cast<const Tin*>(&raw);
            const T w = f.w[k];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const T x = (T)v[j];
                acc[j] = fma(w, x, acc[j]);
                if (k == r) c[j] = x;
            }
        }
        store_vec<VEC>(ic + i * VEC, c);
        store_vec<VEC>(dt0 + i * VEC, acc);
    }
}

// The temporal stage of a BATCH of consecutive output timepoints of a 2D time-lapse in one launch (of3d_flow2d_batch):
// output timepoint j reads the frames j .. j + KT - 1, so a thread keeps a sliding window of KT 16-byte pieces in registers
// (rotated statically over an unrolled period of KT timepoints) and loads every frame ONCE -- 2 (n_out + KT - 1) / n_out
// bytes per pixel instead of 2 KT, and one launch instead of n_out.  8/16-bit frames, paired form only: the arithmetic of
// temporal_vec, operation for operation (bit-identical results).  frames.p[k] = frame k of the batch; ic / dt0 hold n_out
// planes of `plane` elements.
template <typename T, int R>
struct TapsHalf { T w[R + 1]; };       // w[l] = T[R + l]

template <typename Tin, typename T, int KT>
__global__ void __launch_bounds__(256) temporal_vec_batch(const FramePtrs frames, const TapsHalf<T, KT / 2> tw, T* __restrict__ ic,
                                                          T* __restrict__ dt0, int64_t nvec, int n_out, int64_t plane) {
    constexpr int VEC = 16 / sizeof(Tin), R = KT / 2;
    static_assert(std::is_integral<Tin>::value && sizeof(Tin) <= 2, "paired integer form");
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 win[KT];                                      // frame j + k of output timepoint j = j0 + u lives in win[(u + k) % KT]
#pragma unroll
        for (int k = 0; k < KT - 1; ++k) win[k] = __ldg(reinterpret_cast<const uint4*>(frames.p[k]) + i);
#pragma unroll 1
        for (int j0 = 0; j0 < n_out; j0 += KT) {
#pragma unroll
            for (int u = 0; u < KT; ++u) {
                const int j = j0 + u;
                if (j < n_out) {
                    win[(u + KT - 1) % KT] = __ldg(reinterpret_cast<const uint4*>(frames.p[j + KT - 1]) + i);
                    T acc[VEC], c[VEC];
                    const Tin* vc = reinterpret_cast<const Tin*>(&win[(u + R) % KT]);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) { c[q] = (T)vc[q]; acc[q] = T(0); }
#pragma unroll
                    for (int l = 1; l <= R; ++l) {
                        const Tin* va = reinterpret_cast<const Tin*>(&win[(u + R + l) % KT]);
                        const Tin* vb = reinterpret_cast<const Tin*>(&win[(u + R - l) % KT]);
                        const T w = tw.w[l];
#pragma unroll
                        for (int q = 0; q < VEC; ++q) acc[q] = fma(w, (T)((int)va[q] - (int)vb[q]), acc[q]);
                    }
                    store_vec<VEC>(ic + (int64_t)j * plane + i * VEC, c);
                    store_vec<VEC>(dt0 + (int64_t)j * plane + i * VEC, acc);
                }
            }
        }
    }
}

// raw halo planes of the centre frame -> the compute type (what the temporal kernels store as ic: an exact widening)
template <typename Tin, typename T>
__global__ void __launch_bounds__(256) widen_planes(const Tin* __restrict__ in, T* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (T)in[i];
}

// float64 reliability -> the float32 the reference returns in 3D (calc_flow.py:355-357), one rounding (OF3D_FLAG_REL_F32)
__global__ void __launch_bounds__(256) narrow_f64_f32(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}

template <typename T>
__global__ void __launch_bounds__(256) product_generic(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                       int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a[i] * b[i];  // a single multiply: identical to NumPy's in either mode
}

// Window sums, channel-major: w[c*n + i], c in {xx,xy,xz,yy,yz,zz,tx,ty,tz} (3D) / {xx,xy,yy,tx,ty} (2D)
template <typename T, bool EXACT>
__global__ void __launch_bounds__(256) solve3_generic(const T* __restrict__ w, int64_t n, T* __restrict__ vx, T* __restrict__ vy,
                                                      T* __restrict__ vz, T* __restrict__ rel) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Flow3 o = solve3<EXACT>((double)w[i], (double)w[n + i], (double)w[2 * n + i], (double)w[3 * n + i],
                                      (double)w[4 * n + i], (double)w[5 * n + i], (double)w[6 * n + i],
                                      (double)w[7 * n + i], (double)w[8 * n + i]);
        vx[i] = (T)o.vx; vy[i] = (T)o.vy; vz[i] = (T)o.vz; rel[i] = (T)o.rel;
    }
}

template <typename T, bool EXACT>
__global__ void __launch_bounds__(256) solve2_generic(const T* __restrict__ w, int64_t n, T* __restrict__ vx, T* __restrict__ vy,
                                                      T* __restrict__ rel) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Flow2 o = solve2<EXACT>((double)w[i], (double)w[n + i], (double)w[2 * n + i], (double)w[3 * n + i],
                                      (double)w[4 * n + i]);
        vx[i] = (T)o.vx; vy[i] = (T)o.vy; rel[i] = (T)o.rel;
    }
}

}  // namespace of3d
