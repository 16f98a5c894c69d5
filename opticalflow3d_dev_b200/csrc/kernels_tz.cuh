// Temporal derivative + z pass of the gradient stage in ONE marching kernel for sm_100a.  Internal header.
//
// The gradient stage runs z FIRST (separable filters and per-axis clamp-to-edge commute; the reference's order is
// y -> x -> z, calc_flow.py:279-288): with A = S_z Ic, Bz = D_z Ic, C = G_z dt0 the four gradients are
//     dt = G_y G_x C      dy = D_y S_x A      dx = S_y D_x A      dz = S_y S_x Bz
// so the z pass needs three filters instead of four (A is shared by dx and dy) and, marching along z with the lanes on
// x, it can read the RAW frames: a warp stages its 32 columns of all kt frames of a plane with 16-byte cp.async, forms
// dt0 = sum_l T[l] (I[c+l] - I[c-l]) (calc_flow.py:276-278; the paired form is scipy's own evaluation of an antisymmetric
// filter and the differences of 8/16-bit integers are exact) and the widened centre value, and scatters them into three
// statically rotated accumulator rings (kernels_march.cuh).  Neither the widened centre frame nor dt0 ever exists in
// HBM: 2 kt bytes read and 3 values written per voxel (uint16 input).
//
// The march works in batches of ZB = 8 planes.  Per batch: (convert) a COMPACT loop turns the staged raw rows into
// the batch's (Ic, dt0) values in a small per-warp buffer; (issue) the cp.async of the batch after next go into the raw
// buffer just consumed; (scatter) the fully unrolled ring pushes.  Only the scatter is unrolled over the period of the
// rings: with the temporal loop inside every unrolled step the kernel was 160 KB of code and stalled 5 cycles per issue
// on instruction fetch (the L1.5 instruction cache holds 32 KB).
//
// TzSrcPre feeds the same march from (ic, dt0) volumes of the compute type: the two-stage entry points
// (of3d_temporal / of3d_flow_from_dt), input dtypes without a fused instantiation, unaligned rows.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "kernels_march.cuh"

namespace of3d {

constexpr int kTzMaxFrames = 24;      // frames of a fused temporal window (tSig <= 3.6)
constexpr int kTzBatch = 8;           // planes per batch; two batches are in flight
constexpr int kTzWarps = 4;

template <typename T, int KR, int KS>
struct TzArgs {
    MarchGeom g;                      // lanes on x, march z; outputs [g.m_begin, g.m_end)
    Taps<T, KR> fG, fD;
    Taps<T, KS> fS;
    const void* src[kTzMaxFrames];    // RAW: frame k = time c - kt/2 + k; PRE: src[0] = ic, src[1] = dt0 (compute type)
    T tt[kTzMaxFrames / 2 + 1];       // tt[l] = T[kt/2 + l], l = 1 .. kt/2 (T antisymmetric: T[kt/2 - l] = -tt[l])
    int kt;
    int smem_per_warp;                // bytes of shared memory of one warp (Src::smem_per_warp(kt))
    T* out[3];                        // {C = G_z dt0, A = S_z ic, Bz = D_z ic}: (m_end - m_begin) planes, plane 0 = m_begin
};

// exact int -> compute type (|d| < 2^31).  fp64: the magic-number subtraction runs at DFMA rate, the I2F.F64 conversion
// instruction at a quarter of it: 2^52 + 2^31 + d has the low word d ^ 0x80000000.
template <typename T>
__device__ __forceinline__ T tz_from_int(int d) {
    if constexpr (sizeof(T) == 8) return __hiloint2double(0x43300000, d ^ (int)0x80000000) - 4503601774854144.0;
    else return (T)d;
}

// ---- raw frames ------------------------------------------------------------------------------------
// Shared memory per warp: raw[2][ZB][kt][32 * sizeof(Tin)] bytes + vbuf[ZB][2][32] values.  A plane's kt row segments are
// kt * CPR 16-byte pieces; piece p (frame p / CPR, chunk p % CPR) is fetched by lane p % 32 as its (p / 32)-th piece.
// Needs 16-byte aligned frames and rows (checked by the host).  RT = kt / 2 at compile time (0: run-time loop).
template <typename Tin, typename T, int RT>
struct TzSrcRaw {
    static_assert(std::is_integral<Tin>::value && sizeof(Tin) <= 2, "fused temporal stage: 8/16-bit integer frames");
    static constexpr int ZB = kTzBatch;
    static constexpr int ROW = 32 * (int)sizeof(Tin);
    static constexpr int CPR = ROW / 16;
    static constexpr int NP = ((RT ? 2 * RT + 1 : kTzMaxFrames) * CPR + 31) / 32;
    const char* gp[NP];
    uint32_t sbase;            // shared byte address of this lane's piece 0 in buffer 0, plane 0
    const char* lbase;         // generic address of this lane's element in buffer 0, plane 0, frame 0
    T* vbuf;                   // this lane's column of the value buffer
    int slot_bytes, np, kt, rt;
    int qpos, n_march_m1;
    int64_t stride_bytes;

    static size_t smem_per_warp(int kt) { return (size_t)2 * ZB * kt * ROW + (size_t)ZB * 2 * 32 * sizeof(T); }

    template <int KR, int KS>
    __device__ __forceinline__ void init(const TzArgs<T, KR, KS>& a, unsigned char* smem_warp, int64_t lane0, int64_t other, int z_first) {
        const int lane = threadIdx.x & 31;
        const MarchGeom& g = a.g;
        kt = RT ? 2 * RT + 1 : a.kt; rt = kt / 2;
        slot_bytes = kt * ROW;
        const int pieces = kt * CPR;
        np = (pieces + 31) / 32;
        sbase = (uint32_t)__cvta_generic_to_shared(smem_warp) + 16u * lane;
        lbase = reinterpret_cast<const char*>(smem_warp) + lane * sizeof(Tin);
        vbuf = reinterpret_cast<T*>(smem_warp + 2 * ZB * slot_bytes) + lane;
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
    // stage the ZB planes of batch b into raw buffer b % 2 (one cp.async group)
    __device__ __forceinline__ void issue_batch(const int b) {
        const int lane = threadIdx.x & 31;
        uint32_t dst = sbase + (uint32_t)((b & 1) * ZB * slot_bytes);
#pragma unroll 1
        for (int r = 0; r < ZB; ++r) {
            const bool adv = (unsigned)qpos < (unsigned)n_march_m1;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                if (j < np) {
                    if (lane + 32 * j < kt * CPR)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512u * j), "l"(gp[j]) : "memory");
                    if (adv) gp[j] += stride_bytes;
                }
            }
            dst += slot_bytes;
            ++qpos;
        }
        cp_async_commit();
    }
    // raw rows of batch b -> (centre value, temporal derivative) of its ZB planes, in vbuf
    __device__ __forceinline__ void convert(const int b, const T* tt) {
        const char* s = lbase + (b & 1) * ZB * slot_bytes + rt * ROW;          // centre frame of plane 0
#pragma unroll 2
        for (int r = 0; r < ZB; ++r) {
            const T vi = tz_from_int<T>((int)*reinterpret_cast<const Tin*>(s));
            T acc = T(0);
            if constexpr (RT > 0) {
#pragma unroll
                for (int l = 1; l <= RT; ++l) {
                    const int d = (int)*reinterpret_cast<const Tin*>(s + l * ROW) - (int)*reinterpret_cast<const Tin*>(s - l * ROW);
                    acc = fma(tt[l], tz_from_int<T>(d), acc);
                }
            } else {
#pragma unroll 3                                                                 // (loads issued three taps ahead; the FMA chain keeps its order)
                for (int l = 1; l <= rt; ++l) {
                    const int d = (int)*reinterpret_cast<const Tin*>(s + l * ROW) - (int)*reinterpret_cast<const Tin*>(s - l * ROW);
                    acc = fma(tt[l], tz_from_int<T>(d), acc);
                }
            }
            vbuf[(2 * r) * 32] = vi;
            vbuf[(2 * r + 1) * 32] = acc;
            s += slot_bytes;
        }
    }
    __device__ __forceinline__ void value(const int, const int r, T& vi, T& vd) const { vi = vbuf[(2 * r) * 32]; vd = vbuf[(2 * r + 1) * 32]; }
    static constexpr bool kCrossLane = true;     // a lane reads bytes other lanes fetched: __syncwarp around the raw buffers
};

// ---- (ic, dt0) volumes of the compute type -------------------------------------------------------------
// Shared memory per warp: ring[2][ZB][2][32] values, every lane fetches and reads its own column.
template <typename T>
struct TzSrcPre {
    static constexpr int ZB = kTzBatch;
    const char* gp[2];
    uint32_t sbase;
    const T* lbase;
    int qpos, n_march_m1;
    int64_t stride_bytes;
    static size_t smem_per_warp(int) { return (size_t)2 * ZB * 2 * 32 * sizeof(T); }

    template <int KR, int KS>
    __device__ __forceinline__ void init(const TzArgs<T, KR, KS>& a, unsigned char* smem_warp, int64_t lane0, int64_t other, int z_first) {
        const int lane = threadIdx.x & 31;
        const MarchGeom& g = a.g;
        const int64_t lpos = min(lane0 + lane, g.n_lane - 1);
        lbase = reinterpret_cast<const T*>(smem_warp) + lane;
        sbase = (uint32_t)__cvta_generic_to_shared(lbase);
        qpos = z_first;
        n_march_m1 = (int)g.n_march - 1;
        stride_bytes = g.stride_march * (int64_t)sizeof(T);
        const int64_t off0 = other * g.stride_other + lpos + clampi(z_first, g.n_march) * g.stride_march;
        gp[0] = reinterpret_cast<const char*>(static_cast<const T*>(a.src[0]) + off0);
        gp[1] = reinterpret_cast<const char*>(static_cast<const T*>(a.src[1]) + off0);
    }
    __device__ __forceinline__ void issue_batch(const int b) {
        uint32_t dst = sbase + (uint32_t)((b & 1) * ZB * 2 * 32 * sizeof(T));
#pragma unroll 1
        for (int r = 0; r < ZB; ++r) {
            cp_async_elem<T>(dst, gp[0]);
            cp_async_elem<T>(dst + (uint32_t)(32 * sizeof(T)), gp[1]);
            if ((unsigned)qpos < (unsigned)n_march_m1) { gp[0] += stride_bytes; gp[1] += stride_bytes; }
            dst += (uint32_t)(2 * 32 * sizeof(T));
            ++qpos;
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void convert(const int, const T*) {}
    __device__ __forceinline__ void value(const int b, const int r, T& vi, T& vd) const {
        const T* s = lbase + ((b & 1) * ZB + r) * 2 * 32;
        vi = s[0]; vd = s[32];
    }
    static constexpr bool kCrossLane = false;
};

// One warp per (y row, 32 x columns, z chunk).  Step s consumes input plane c0 - R + s (clamped to the volume): the
// G and D rings (radius R) complete output c0 + s - 2R, the S ring (radius RS) output c0 + s - R - RS.
// SHIFT: the three marches run on shifting rings (kernels_march.cuh shift_push): 2 KR + KS accumulators instead of 3 P, one
// copy of a batch in the instruction stream.
template <typename Src, typename T, int KR, int KS, int P, int WPB, bool SHIFT = false, int MINB = 0>
__global__ void __launch_bounds__(WPB * 32, MINB ? MINB : ((sizeof(T) == 8 && P > 24) ? 2 : 3)) march_tz(const TzArgs<T, KR, KS> a) {
    constexpr int R = KR / 2, RS = KS / 2, ZB = kTzBatch;
    static_assert(P >= KR && KR >= KS && P % ZB == 0, "bad unroll period");
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
    src.issue_batch(0);
    src.issue_batch(1);

    // the three output pointers of the plane the R rings complete at step s (plane 0 of the outputs is m_begin); the S
    // ring runs R - RS planes ahead
    const int64_t ostride = g.stride_march * (int64_t)sizeof(T);
    const int64_t o0 = (other * g.stride_other + min(lane0 + lane, g.n_lane - 1)) * (int64_t)sizeof(T) + ((int64_t)(c0 - g.m_begin) - 2 * R) * ostride;
    char* pG = reinterpret_cast<char*>(a.out[0]) + o0;
    char* pD = reinterpret_cast<char*>(a.out[2]) + o0;
    char* pS = reinterpret_cast<char*>(a.out[1]) + o0 + (int64_t)(R - RS) * ostride;

    if constexpr (SHIFT) {
        T sG[KR], sD[KR], sS[KS];
#pragma unroll
        for (int i = 0; i < KR; ++i) { sG[i] = T(0); sD[i] = T(0); }
#pragma unroll
        for (int i = 0; i < KS; ++i) sS[i] = T(0);
        const int nb = (nout + 2 * R + ZB - 1) / ZB;
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            cp_async_wait<1>();
            if (Src::kCrossLane) __syncwarp();
            src.convert(b, a.tt);
            if (Src::kCrossLane) __syncwarp();
            if (Src::kCrossLane) src.issue_batch(b + 2);
#pragma unroll
            for (int r = 0; r < ZB; ++r) {
                const int s = b * ZB + r;
                T vi, vd;
                src.value(b, r, vi, vd);
                const T rG = shift_push<T, KR, 1>(sG, a.fG, vd);
                const T rD = shift_push<T, KR, -1>(sD, a.fD, vi);
                const T rS = shift_push<T, KS, 1>(sS, a.fS, vi);
                if ((unsigned)(s - 2 * R) < nvalid) {
                    *reinterpret_cast<T*>(pG) = rG;
                    *reinterpret_cast<T*>(pD) = rD;
                }
                if ((unsigned)(s - R - RS) < nvalid) *reinterpret_cast<T*>(pS) = rS;
                pG += ostride; pD += ostride; pS += ostride;
            }
            if (!Src::kCrossLane) src.issue_batch(b + 2);
        }
        cp_async_wait<0>();
        return;
    }
    T accG[P], accD[P], accS[P];
#pragma unroll
    for (int i = 0; i < P; ++i) { accG[i] = T(0); accD[i] = T(0); accS[i] = T(0); }
    int b = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int sg = 0; sg < P / ZB; ++sg, ++b) {
            cp_async_wait<1>();                    // batch b has landed, b + 1 is in flight
            if (Src::kCrossLane) __syncwarp();
            src.convert(b, a.tt);
            if (Src::kCrossLane) __syncwarp();     // every lane is done with the raw rows of batch b
            if (Src::kCrossLane) src.issue_batch(b + 2);
#pragma unroll
            for (int r = 0; r < ZB; ++r) {
                const int ph = sg * ZB + r;
                T vi, vd;
                src.value(b, r, vi, vd);
                const T rG = ring_push<T, KR, P, 1>(accG, a.fG, vd, ph);
                const T rD = ring_push<T, KR, P, -1>(accD, a.fD, vi, ph);
                const T rS = ring_push<T, KS, P, 1>(accS, a.fS, vi, ph);
                if ((unsigned)(s0 + ph - 2 * R) < nvalid) {
                    *reinterpret_cast<T*>(pG) = rG;
                    *reinterpret_cast<T*>(pD) = rD;
                }
                if ((unsigned)(s0 + ph - R - RS) < nvalid) *reinterpret_cast<T*>(pS) = rS;
                pG += ostride; pD += ostride; pS += ostride;
            }
            if (!Src::kCrossLane) src.issue_batch(b + 2);   // the ring itself was the value buffer: refill it after the scatter
        }
    }
    cp_async_wait<0>();
}

}  // namespace of3d
