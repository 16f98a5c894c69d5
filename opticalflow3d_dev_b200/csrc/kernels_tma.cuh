// TMA-fed marching kernels for sm_100a.  Internal header.
//
// The cp.async (LDGSTS) marches of kernels_march.cuh pay 8 LSU cycles per 32-lane load instruction and every channel
// re-fetches the gradient rows it shares with other channels: the products + window-z march ran at 83 % L1/LSU
// utilisation with the FP64 pipe 62 % busy.  Here one block owns (y row, 64 x columns) for ALL nine channels: the
// tile {64 x, 1 y, ZT z, 4 gradient volumes} of a stage is fetched ONCE per block by one tensor-map copy
// (cp.async.bulk.tensor, the TMA engine: no LSU instruction per lane, no L1 wavefronts) into a shared-memory ring of
// STAGES stages, completion tracked by an mbarrier transaction count; the channel warps wait on the stage's "full"
// barrier, march ZT steps out of shared memory and release the stage on its "empty" barrier.
// TMA fills out-of-range elements with zeros, scipy's mode='nearest' wants the edge plane: a stage that reaches beyond
// a z face is fetched plane by plane (ZT copies of a one-plane box, z clamped) instead of as one ZT-plane box.  The
// tensor map orders the dimensions (x, y, volume, z), so that a stage is [z][volume][x] in shared memory and a
// one-plane box lands on one contiguous z slice of it.
// (Row-by-row cp.async.bulk copies with clamped addresses were tried first: 64 copies of 256 B per stage ran at
// ~80 cycles per copy, 0.9 TB/s over the chip.)
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "kernels_march.cuh"

namespace of3d {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// one 4-D tile global -> shared through a tensor map, completion (box bytes) on `bar`
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int x, int y, int z, int v, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(v), "r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Products + window z pass (calc_flow.py:300-313, z passes), all nine channels of a (y, 32 NG x) column per block:
// 9 NG marching warps (channel ch = warp % 9 of 32-column group warp / 9) + one producer warp.  The producer loop lives in
// a __noinline__ function: inlined anywhere in the kernel it makes ptxas move the taps from uniform registers to 50
// vector registers (168 registers unbounded, spills at the 96 the block size allows).
// NG = 2 (19 warps, 104 registers) up to 25 taps in fp64 and for every fp32 window; NG = 1 (10 warps) for the long fp64
// windows whose accumulator ring alone needs 80-112 registers.
constexpr int kTmaZT = 8;        // z steps per stage
constexpr int kTmaStages = 4;
constexpr int kEdgeGroup = 8;    // taps per uniform branch at the tail of a march (ring_push_tail)

template <typename T, int NG>
constexpr size_t window_tma_smem() { return (size_t)kTmaStages * 4 * kTmaZT * 32 * NG * sizeof(T) + 2 * kTmaStages * 8; }

// One block per (y, column group, z chunk) task.  (Persistent blocks striding over the tasks, so that the ring never
// drains between tasks, and a 6-stage ring were measured 3 % slower: the task bookkeeping costs registers the marching
// loop spills for.)
struct TmaTask {
    int x0, y, c0, nout, nstages;
};
template <int K, int NG>
__device__ __forceinline__ TmaTask tma_task(int n_chunks, int lane_groups, int chunk_len, int m_begin, int m_end, int task) {
    TmaTask t;
    const int chunk = task % n_chunks; task /= n_chunks;
    const int ngrp = (lane_groups + NG - 1) / NG;
    t.x0 = (task % ngrp) * 32 * NG;
    t.y = task / ngrp;
    t.c0 = m_begin + chunk * chunk_len;
    t.nout = min(t.c0 + chunk_len, m_end) - t.c0;
    t.nstages = (t.nout + 2 * (K / 2) + kTmaZT - 1) / kTmaZT;
    return t;
}

// tm8: box {32 NG, 1, 4, ZT}; tm1: box {32 NG, 1, 4, 1} over the same tensor
template <typename T, int K, int NG>
__device__ __noinline__ void window_tma_producer(const CUtensorMap* tm8, const CUtensorMap* tm1, int n_chunks, int lane_groups, int chunk_len,
                                                 int m_begin, int m_end, int nz, uint32_t ring_s, uint32_t bar_s) {
    constexpr int ZT = kTmaZT, ST = kTmaStages, R = K / 2;
    constexpr uint32_t kPlaneBytes = 4 * 32 * NG * sizeof(T);
    constexpr uint32_t kStageBytes = ZT * kPlaneBytes;
    if ((threadIdx.x & 31) != 0) return;
    const TmaTask t = tma_task<K, NG>(n_chunks, lane_groups, chunk_len, m_begin, m_end, blockIdx.x);
    for (int it = 0; it < t.nstages; ++it) {
        const int slot = it % ST;
        if (it >= ST) mbar_wait(bar_s + 8 * (ST + slot), ((it / ST) - 1) & 1);       // every marching warp released the slot
        mbar_arrive_expect_tx(bar_s + 8 * slot, kStageBytes);
        const int z0 = t.c0 - R + it * ZT;                                            // first input plane of the stage
        if (z0 >= 0 && z0 + ZT <= nz) {
            tma_load_4d(ring_s + slot * kStageBytes, tm8, t.x0, t.y, 0, z0, bar_s + 8 * slot);
        } else {
            for (int r = 0; r < ZT; ++r)                                               // clamp-to-edge (scipy mode='nearest')
                tma_load_4d(ring_s + slot * kStageBytes + r * kPlaneBytes, tm1, t.x0, t.y, 0, max(0, min(z0 + r, nz - 1)), bar_s + 8 * slot);
        }
    }
}

// tensor maps: {x, y, 4 volumes, z} over the gradient volumes {dt, dx, dy, dz}
template <typename T, int K, int P, int NG, int MINB = 1>
__global__ void __launch_bounds__((9 * NG + 1) * 32, MINB) march_window_tma(const WindowArgs<T, K> a, const __grid_constant__ CUtensorMap tm8,
                                                                          const __grid_constant__ CUtensorMap tm1) {
    static_assert(P >= K && P % kTmaZT == 0, "bad unroll period");
    constexpr int R = K / 2, ZT = kTmaZT, ST = kTmaStages, TX = 32 * NG;
    constexpr int kStageElems = 4 * ZT * TX;                 // [z][volume][x]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* ring = reinterpret_cast<T*>(smem_raw);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t bar_s = ring_s + (uint32_t)(ST * kStageElems * sizeof(T));   // full[ST], empty[ST]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const MarchGeom& g = a.g;

    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; ++i) { mbar_init(bar_s + 8 * i, 1); mbar_init(bar_s + 8 * (ST + i), 9 * NG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 9 * NG) {
        window_tma_producer<T, K, NG>(&tm8, &tm1, g.n_chunks, g.lane_groups, g.chunk, g.m_begin, g.m_end, (int)g.n_march, ring_s, bar_s);
        return;
    }

    // ---- marching warps
    const int ch = warp % 9, grp = warp / 9;
    // channel -> gradient pair {xx,xy,xz,yy,yz,zz,tx,ty,tz} over {dt,dx,dy,dz}
    const int ia = ch < 3 ? 1 : (ch < 5 ? 2 : (ch == 5 ? 3 : ch - 5));
    const int ib = ch < 3 ? ch + 1 : (ch < 5 ? ch - 1 : (ch == 5 ? 3 : 0));
    const T* sa = ring + ia * TX + grp * 32 + lane;
    const T* sb = ring + ib * TX + grp * 32 + lane;
    const int64_t stride_bytes = g.stride_march * (int64_t)sizeof(T);

    T acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) acc[i] = T(0);

    const TmaTask t = tma_task<K, NG>(g.n_chunks, g.lane_groups, g.chunk, g.m_begin, g.m_end, blockIdx.x);
    const int64_t lane0 = (int64_t)t.x0 + grp * 32;
    const unsigned nvalid = lane0 + lane < g.n_lane ? (unsigned)t.nout : 0u;
    // store position of the output completed at step s: c0 + s - 2R (plane 0 of the output is m_begin)
    char* optr = reinterpret_cast<char*>(a.out + (int64_t)ch * g.vol + (int64_t)t.y * g.stride_other + lane0 + lane) +
                 ((int64_t)(t.c0 - g.m_begin) - 2 * R) * stride_bytes;
    int it = 0;
    // first period: the warm-up triangle of the ring is skipped (ring_push<..., WARM>)
#pragma unroll
    for (int sg = 0; sg < P / ZT; ++sg) {
        const int slot = it % ST;
        mbar_wait(bar_s + 8 * slot, (it / ST) & 1);
        const T* pa = sa + slot * kStageElems;
        const T* pb = sb + slot * kStageElems;
#pragma unroll
        for (int r = 0; r < ZT; ++r) {
            const int ph = sg * ZT + r;
            const T v = pa[r * 4 * TX] * pb[r * 4 * TX];
            const T res = ring_push<T, K, P, 1, true>(acc, a.f, v, ph);
            if ((unsigned)(ph - 2 * R) < nvalid) *reinterpret_cast<T*>(optr) = res;
            optr += stride_bytes;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_s + 8 * (ST + slot));
        if (++it >= t.nstages) return;
    }
#pragma unroll 1
    for (int s0 = P;; s0 += P) {
        // step s of the march lies d = s - nout steps into the tail: its taps k <= d only feed outputs beyond the range.
        // A stage whose first step has d > 0 runs the copy of the body that skips them in groups (ring_push_tail).
        // edge_skip == 0 (OF3D_NO_TAIL_SKIP): never in the tail
        const int dbase = a.edge_skip ? s0 - t.nout : -(1 << 20);
#pragma unroll
        for (int sg = 0; sg < P / ZT; ++sg) {
            const int slot = it % ST;
            mbar_wait(bar_s + 8 * slot, (it / ST) & 1);
            const T* pa = sa + slot * kStageElems;
            const T* pb = sb + slot * kStageElems;
            const int d0 = dbase + sg * ZT;
            // (windows longer than 25 taps: one copy of the body, the group tests always run -- two copies of a 56 x 49
            // scatter are more than the compiler unrolls, and the accumulator ring ends up in local memory)
            if (K <= 25 && d0 <= 0) {
#pragma unroll
                for (int r = 0; r < ZT; ++r) {
                    const int ph = sg * ZT + r;
                    const T v = pa[r * 4 * TX] * pb[r * 4 * TX];
                    const T res = ring_push<T, K, P, 1>(acc, a.f, v, ph);
                    if ((unsigned)(s0 + ph - 2 * R) < nvalid) *reinterpret_cast<T*>(optr) = res;
                    optr += stride_bytes;
                }
            } else {
#pragma unroll
                for (int r = 0; r < ZT; ++r) {
                    const int ph = sg * ZT + r;
                    const T v = pa[r * 4 * TX] * pb[r * 4 * TX];
                    const T res = ring_push_tail<T, K, P, 1, kEdgeGroup>(acc, a.f, v, ph, d0);
                    if ((unsigned)(s0 + ph - 2 * R) < nvalid) *reinterpret_cast<T*>(optr) = res;
                    optr += stride_bytes;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_s + 8 * (ST + slot));
            if (++it >= t.nstages) return;
        }
    }
}

// The same kernel on the SHIFTING ring (kernels_march.cuh shift_push): one loop over the stages, the ZT steps of a stage
// unrolled, K accumulators.  Interior stages run the plain step; the stages that contain warm-up or tail steps run the
// copy with the group tests.
template <typename T, int K, int NG, int MINB = 1>
__global__ void __launch_bounds__((9 * NG + 1) * 32, MINB) march_window_tma_sh(const WindowArgs<T, K> a, const __grid_constant__ CUtensorMap tm8,
                                                                             const __grid_constant__ CUtensorMap tm1) {
    constexpr int R = K / 2, ZT = kTmaZT, ST = kTmaStages, TX = 32 * NG;
    constexpr int kStageElems = 4 * ZT * TX;                 // [z][volume][x]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* ring = reinterpret_cast<T*>(smem_raw);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t bar_s = ring_s + (uint32_t)(ST * kStageElems * sizeof(T));   // full[ST], empty[ST]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const MarchGeom& g = a.g;

    if (threadIdx.x == 0) {
        for (int i = 0; i < ST; ++i) { mbar_init(bar_s + 8 * i, 1); mbar_init(bar_s + 8 * (ST + i), 9 * NG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 9 * NG) {
        window_tma_producer<T, K, NG>(&tm8, &tm1, g.n_chunks, g.lane_groups, g.chunk, g.m_begin, g.m_end, (int)g.n_march, ring_s, bar_s);
        return;
    }

    const int ch = warp % 9, grp = warp / 9;
    const int ia = ch < 3 ? 1 : (ch < 5 ? 2 : (ch == 5 ? 3 : ch - 5));
    const int ib = ch < 3 ? ch + 1 : (ch < 5 ? ch - 1 : (ch == 5 ? 3 : 0));
    const T* sa = ring + ia * TX + grp * 32 + lane;
    const T* sb = ring + ib * TX + grp * 32 + lane;
    const int64_t stride_bytes = g.stride_march * (int64_t)sizeof(T);

    T acc[K];
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] = T(0);

    const TmaTask t = tma_task<K, NG>(g.n_chunks, g.lane_groups, g.chunk, g.m_begin, g.m_end, blockIdx.x);
    const int64_t lane0 = (int64_t)t.x0 + grp * 32;
    const unsigned nvalid = lane0 + lane < g.n_lane ? (unsigned)t.nout : 0u;
    // store position of the output completed at step s: c0 + s - 2R (plane 0 of the output is m_begin)
    char* optr = reinterpret_cast<char*>(a.out + (int64_t)ch * g.vol + (int64_t)t.y * g.stride_other + lane0 + lane) +
                 ((int64_t)(t.c0 - g.m_begin) - 2 * R) * stride_bytes;
    const int tail0 = a.edge_skip ? t.nout : (1 << 28);       // first step of the tail
#pragma unroll 1
    for (int it = 0; it < t.nstages; ++it) {
        const int slot = it % ST;
        mbar_wait(bar_s + 8 * slot, (it / ST) & 1);
        const T* pa = sa + slot * kStageElems;
        const T* pb = sb + slot * kStageElems;
        const int s0 = it * ZT;
        if (s0 >= 2 * R && s0 + ZT - 1 <= tail0) {
#pragma unroll
            for (int r = 0; r < ZT; ++r) {
                const T v = pa[r * 4 * TX] * pb[r * 4 * TX];
                const T res = shift_push<T, K, 1>(acc, a.f, v);
                if ((unsigned)(s0 + r - 2 * R) < nvalid) *reinterpret_cast<T*>(optr) = res;
                optr += stride_bytes;
            }
        } else {
#pragma unroll
            for (int r = 0; r < ZT; ++r) {
                const T v = pa[r * 4 * TX] * pb[r * 4 * TX];
                // (one set of group tests per stage: its last step decides the warm-up taps, its first step the tail taps)
                const T res = shift_push_edge<T, K, 1, kEdgeGroup>(acc, a.f, v, s0 + ZT - 1, s0 - tail0);
                if ((unsigned)(s0 + r - 2 * R) < nvalid) *reinterpret_cast<T*>(optr) = res;
                optr += stride_bytes;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_s + 8 * (ST + slot));
    }
}

}  // namespace of3d
