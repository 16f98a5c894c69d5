// Temporal derivative + z pass of the gradient stage in ONE marching kernel for sm_100a.  Internal header.
//
// The gradient stage runs z FIRST (separable filters and per-axis clamp-to-edge commute; the reference's order is
// y -> x -> z, calc_flow.py:279-288): with A = S_z Ic, Bz = D_z Ic, C = G_z dt0 the four gradients are
//     dt = G_y G_x C      dy = D_y S_x A      dx = S_y D_x A      dz = S_y S_x Bz
// so the z pass needs three filters instead of four (A is shared by dx and dy) and, marching along z with the lanes on
// x, it can read the RAW frames: a warp stages its 32 columns of all kt frames of a plane with 16-byte cp.async, forms
// dt0 = sum_l T[l] (I[c+l] - I[c-l]) (calc_flow.py:276-278; the paired form is scipy's own evaluation of an antisymmetric
// filter and the differences of 8/16-bit integers are exact) and the widened centre value in registers, and scatters
// them into three statically rotated accumulator rings (kernels_march.cuh).  Neither the widened centre frame nor dt0
// ever exists in HBM: 2 kt bytes read and 3 values written per voxel (uint16 input).
//
// TzSrcPre feeds the same march from (ic, dt0) volumes of the compute type: the two-stage entry points
// (of3d_temporal / of3d_flow_from_dt), input dtypes without a fused instantiation, unaligned rows.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "kernels_march.cuh"

namespace of3d {

constexpr int kTzMaxFrames = 24;      // frames of a fused temporal window (tSig <= 3.6)
constexpr int kTzDepth = 8;           // z steps in flight per warp
constexpr int kTzWarps = 4;

template <typename T, int KR, int KS>
struct TzArgs {
    MarchGeom g;                      // lanes on x, march z; outputs [g.m_begin, g.m_end)
    Taps<T, KR> fG, fD;
    Taps<T, KS> fS;
    const void* src[kTzMaxFrames];    // RAW: frame k = time c - kt/2 + k; PRE: src[0] = ic, src[1] = dt0 (compute type)
    T tt[kTzMaxFrames / 2 + 1];       // tt[l] = T[kt/2 + l], l = 1 .. kt/2 (T antisymmetric: T[kt/2 - l] = -tt[l])
    int kt;
    int smem_per_warp;                // bytes of the prefetch ring of one warp (Src::smem_per_warp(kt))
    T* out[3];                        // {C = G_z dt0, A = S_z ic, Bz = D_z ic}: (m_end - m_begin) planes, plane 0 = m_begin
};

// exact int -> compute type (|d| < 2^24 for the fused dtypes).  fp64: the magic-number add runs at DFMA rate; the
// I2F.F64 conversion instruction has a quarter of that throughput.
template <typename T>
__device__ __forceinline__ T tz_from_int(int d) {
    if constexpr (sizeof(T) == 8) return __hiloint2double(0x43380000, d) - 6755399441055744.0;   // 2^52 + 2^51 + d
    else return (T)d;
}

// ---- raw frames ------------------------------------------------------------------------------------
// Shared-memory ring per warp: [DEPTH][kt][32 * sizeof(Tin)] bytes.  A step's kt row segments are kt * CPR 16-byte
// pieces; piece p (frame p / CPR, chunk p % CPR) is fetched by lane p % 32 as its (p / 32)-th piece.  Needs 16-byte
// aligned frames and rows (checked by the host).
template <typename Tin, typename T, int DEPTH>
struct TzSrcRaw {
    static_assert(std::is_integral<Tin>::value && sizeof(Tin) <= 2, "fused temporal stage: 8/16-bit integer frames");
    static constexpr int ROW = 32 * (int)sizeof(Tin);
    static constexpr int CPR = ROW / 16;
    static constexpr int NP = (kTzMaxFrames * CPR + 31) / 32;
    const char* gp[NP];
    uint32_t sbase;            // shared byte address of this lane's piece 0 in slot 0
    const char* lbase;         // shared address (generic) of this lane's element in slot 0, frame 0
    int slot_bytes, np, kt, rt;
    int qpos, n_march_m1;
    int64_t stride_bytes;

    static size_t smem_per_warp(int kt) { return (size_t)DEPTH * kt * ROW; }

    template <int KR, int KS>
    __device__ __forceinline__ void init(const TzArgs<T, KR, KS>& a, unsigned char* smem_warp, int64_t lane0, int64_t other, int z_first) {
        const int lane = threadIdx.x & 31;
        const MarchGeom& g = a.g;
        kt = a.kt; rt = kt / 2;
        slot_bytes = kt * ROW;
        const int pieces = kt * CPR;
        np = (pieces + 31) / 32;
        sbase = (uint32_t)__cvta_generic_to_shared(smem_warp) + 16u * lane;
        lbase = reinterpret_cast<const char*>(smem_warp) + lane * sizeof(Tin);
        qpos = z_first;
        n_march_m1 = (int)g.n_march - 1;
        stride_bytes = g.stride_march * (int64_t)sizeof(Tin);
        const int64_t zc = clampi(z_first, g.n_march);
        const int64_t row_bytes = g.n_lane * (int64_t)sizeof(Tin);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int p = min(lane + 32 * j, pieces - 1);
            const int k = p / CPR, cx = p % CPR;
            // chunks beyond the end of the row (last lane group of a row that is not a multiple of 32) re-read the last chunk
            const int64_t xb = min(lane0 * (int64_t)sizeof(Tin) + 16 * cx, row_bytes - 16);
            gp[j] = static_cast<const char*>(a.src[k]) + (other * g.stride_other + zc * g.stride_march) * (int64_t)sizeof(Tin) + xb;
        }
    }
    __device__ __forceinline__ void issue(const int slot) {
        const int lane = threadIdx.x & 31;
        const bool adv = (unsigned)qpos < (unsigned)n_march_m1;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            if (j < np) {
                if (lane + 32 * j < kt * CPR)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (uint32_t)(slot * slot_bytes + 512 * j)), "l"(gp[j]) : "memory");
                if (adv) gp[j] += stride_bytes;
            }
        }
        cp_async_commit();
        ++qpos;
    }
    // values of the step in `slot`: centre frame (widened) and temporal derivative
    __device__ __forceinline__ void read(const int slot, const T* tt, T& vi, T& vd) const {
        const char* s = lbase + slot * slot_bytes;
        const Tin* c = reinterpret_cast<const Tin*>(s + rt * ROW);
        vi = tz_from_int<T>((int)c[0]);
        T acc = T(0);
        for (int l = 1; l <= rt; ++l) {
            const int d = (int)*reinterpret_cast<const Tin*>(reinterpret_cast<const char*>(c) + l * ROW) -
                          (int)*reinterpret_cast<const Tin*>(reinterpret_cast<const char*>(c) - l * ROW);
            acc = fma(tt[l], tz_from_int<T>(d), acc);
        }
        vd = acc;
    }
    static constexpr bool kCrossLane = true;     // a lane reads bytes other lanes fetched: __syncwarp after the wait
};

// ---- (ic, dt0) volumes of the compute type -------------------------------------------------------------
template <typename T, int DEPTH>
struct TzSrcPre {
    Prefetcher<T, 2, DEPTH> pre;
    static size_t smem_per_warp(int) { return Prefetcher<T, 2, DEPTH>::elems_per_warp * sizeof(T); }

    template <int KR, int KS>
    __device__ __forceinline__ void init(const TzArgs<T, KR, KS>& a, unsigned char* smem_warp, int64_t lane0, int64_t other, int z_first) {
        const int lane = threadIdx.x & 31;
        const MarchGeom& g = a.g;
        const int64_t lpos = min(lane0 + lane, g.n_lane - 1);
        pre.lbase = reinterpret_cast<T*>(smem_warp) + lane;
        pre.sbase = (uint32_t)__cvta_generic_to_shared(pre.lbase);
        pre.qpos = z_first;
        pre.n_march_m1 = (int)g.n_march - 1;
        pre.stride_bytes = g.stride_march * (int64_t)sizeof(T);
        const int64_t off0 = other * g.stride_other + lpos + clampi(z_first, g.n_march) * g.stride_march;
        pre.gp[0] = reinterpret_cast<const char*>(static_cast<const T*>(a.src[0]) + off0);
        pre.gp[1] = reinterpret_cast<const char*>(static_cast<const T*>(a.src[1]) + off0);
    }
    __device__ __forceinline__ void issue(const int slot) { pre.issue(slot); }
    __device__ __forceinline__ void read(const int slot, const T*, T& vi, T& vd) const { vi = pre.read(slot, 0); vd = pre.read(slot, 1); }
    static constexpr bool kCrossLane = false;
};

// One warp per (y row, 32 x columns, z chunk).  Step s consumes input plane c0 - R + s (clamped to the volume): the
// G and D rings (radius R) complete output c0 + s - 2R, the S ring (radius RS) output c0 + s - R - RS.
template <typename Src, typename T, int KR, int KS, int P, int WPB>
__global__ void __launch_bounds__(WPB * 32) march_tz(const TzArgs<T, KR, KS> a) {
    constexpr int R = KR / 2, RS = KS / 2, DEPTH = kTzDepth;
    static_assert(P >= KR && KR >= KS && P % DEPTH == 0, "bad unroll period");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const MarchGeom& g = a.g;
    int64_t task = (int64_t)blockIdx.x * WPB + warp;
    const int64_t ntasks = (int64_t)g.n_other * g.lane_groups * g.n_chunks;
    if (task >= ntasks) return;
    const int chunk = (int)(task % g.n_chunks); task /= g.n_chunks;
    const int lg = (int)(task % g.lane_groups);
    const int64_t other = task / g.lane_groups;
    const int64_t lane0 = (int64_t)lg * 32;
    const bool lane_ok = lane0 + lane < g.n_lane;
    const int c0 = g.m_begin + chunk * g.chunk;
    const int c1 = min(c0 + g.chunk, g.m_end);
    const int nout = c1 - c0;
    const unsigned nvalid = lane_ok ? (unsigned)nout : 0u;
    const int nsteps = (nout + 2 * R + P - 1) / P * P;

    Src src;
    src.init(a, smem_raw + (size_t)warp * a.smem_per_warp, lane0, other, c0 - R);
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) src.issue(d);

    T accG[P], accD[P], accS[P];
#pragma unroll
    for (int i = 0; i < P; ++i) { accG[i] = T(0); accD[i] = T(0); accS[i] = T(0); }
    // byte offset (in the outputs, whose plane 0 is m_begin) of the output completed by the R rings at step s
    const int64_t ostride = g.stride_march * (int64_t)sizeof(T);
    int64_t ooff = (other * g.stride_other + min(lane0 + lane, g.n_lane - 1)) * (int64_t)sizeof(T) + ((int64_t)(c0 - g.m_begin) - 2 * R) * ostride;
    const int64_t soff = (int64_t)(R - RS) * ostride;       // the S ring runs R - RS planes ahead

    cp_async_wait<DEPTH - 2>();
    if (Src::kCrossLane) __syncwarp();
    T ni, nd;
    src.read(0, a.tt, ni, nd);
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const T vi = ni, vd = nd;
            src.issue((ph + DEPTH - 1) % DEPTH);
            cp_async_wait<DEPTH - 2>();            // step s+1 has landed: read it while step s is accumulated
            if (Src::kCrossLane) __syncwarp();
            src.read((ph + 1) % DEPTH, a.tt, ni, nd);
            const T rG = ring_push<T, KR, P, 1>(accG, a.fG, vd, ph);
            const T rD = ring_push<T, KR, P, -1>(accD, a.fD, vi, ph);
            const T rS = ring_push<T, KS, P, 1>(accS, a.fS, vi, ph);
            if ((unsigned)(s0 + ph - 2 * R) < nvalid) {
                *reinterpret_cast<T*>(reinterpret_cast<char*>(a.out[0]) + ooff) = rG;
                *reinterpret_cast<T*>(reinterpret_cast<char*>(a.out[2]) + ooff) = rD;
            }
            if ((unsigned)(s0 + ph - R - RS) < nvalid) *reinterpret_cast<T*>(reinterpret_cast<char*>(a.out[1]) + ooff + soff) = rS;
            ooff += ostride;
        }
    }
    cp_async_wait<0>();
}

}  // namespace of3d
