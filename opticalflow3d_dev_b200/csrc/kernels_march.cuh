// Marching ("scatter") separable-correlation kernels for sm_100a.  Internal header.
//
// Idea: one warp owns 32 adjacent positions of the CONTIGUOUS axis (the lanes) and marches along the
// filtered axis.  Every input it loads (one coalesced 128/256-byte row segment per step) is multiplied
// into the K pending outputs it contributes to; the K partial sums live in registers, and the ring of
// accumulators is rotated STATICALLY by fully unrolling P >= K steps, so each tap is an immediate
// constant-bank operand of one DFMA/FFMA.  One global load per K FMAs, no halo re-reads along the
// filtered axis.  The accumulators leave room for only 2-4 warps per scheduler, and a warp has just six
// scoreboard slots for register loads, so the inputs are prefetched with cp.async (LDGSTS) into a
// per-warp shared-memory ring DEPTH steps ahead: completion is tracked by cp.async groups, not by the
// scoreboard, which is what hides the ~1 us HBM latency at this occupancy.
//
// Marching needs the lanes on the contiguous axis (x), so it is used for the two z passes (gradient z pass,
// products + window z pass); the in-plane x and y passes are fused in kernels_strip.cuh.  Pass order:
// gradients x -> y -> z, window z -> x -> y; separable filters and per-axis clamp-to-edge commute, so this
// equals the reference's y -> x -> z up to rounding.
#pragma once
#include <cstring>
#include "common.cuh"
#include "solve.cuh"

namespace of3d {

template <typename T, int K>
struct Taps {
    T w[K];
};

// Geometry of one marching pass over a volume with axes (other, march, lane) in arbitrary memory order;
// the lane axis has stride 1.
struct MarchGeom {
    int64_t n_lane;        // extent of the lane (contiguous) axis
    int64_t n_march;       // extent of the filtered axis
    int64_t n_other;       // extent of the remaining axis
    int64_t stride_march;  // element stride of the filtered axis (input and plain output)
    int64_t stride_other;  // element stride of the remaining axis (same for the transposed output)
    int64_t vol;           // elements per volume (channel stride)
    int m_begin, m_end;    // outputs [m_begin, m_end) of the filtered axis are produced (inputs clamp to [0, n_march))
    int chunk;             // outputs per task along the filtered axis (multiple of 32)
    int n_chunks;
    int lane_groups;       // ceil(n_lane / 32)
};

__device__ __forceinline__ int64_t clampi(int64_t p, int64_t n) { return p < 0 ? 0 : (p >= n ? n - 1 : p); }

// cp.async of one element (4 or 8 bytes) per lane; completion by commit/wait groups
template <typename T>
__device__ __forceinline__ void cp_async_elem(uint32_t saddr, const void* gptr) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "element size");
    if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(gptr) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-warp prefetch ring in shared memory, [DEPTH][NIN][32 lanes].  Step q of the march lives in slot
// q % DEPTH; the unroll period of the march is a multiple of DEPTH, so every slot index is a compile-time
// constant and the LDS / LDGSTS addresses are base + immediate.  Each step first issues step s + DEPTH - 1
// (into the slot that was read one step earlier) and then waits for its own data, so DEPTH - 1 steps are
// always in flight.  The global pointers advance by one row per step, except outside [0, n_march) where
// they stay on the edge row: that is scipy's mode='nearest' (clamp to edge).
template <typename T, int NIN, int DEPTH>
struct Prefetcher {
    const char* gp[NIN];   // next row to fetch, per input (already offset to this lane)
    uint32_t sbase;        // shared byte address of this lane in slot 0, input 0
    const T* lbase;        // same location as a pointer (for reads)
    int qpos;              // input position of the next fetch (may be < 0 or >= n_march)
    int n_march_m1;
    int64_t stride_bytes;
    static constexpr int kSlotElems = NIN * 32;
    static constexpr int elems_per_warp = DEPTH * NIN * 32;

    __device__ __forceinline__ void issue(const int slot) {
#pragma unroll
        for (int i = 0; i < NIN; ++i) cp_async_elem<T>(sbase + (uint32_t)((slot * NIN + i) * 32 * sizeof(T)), gp[i]);
        cp_async_commit();
        if ((unsigned)qpos < (unsigned)n_march_m1) {   // 0 <= qpos < n_march - 1: the next row exists
#pragma unroll
            for (int i = 0; i < NIN; ++i) gp[i] += stride_bytes;
        }
        ++qpos;
    }
    __device__ __forceinline__ T read(const int slot, const int i) const { return lbase[(slot * NIN + i) * 32]; }
};

// The scatter FMAs are volatile asm so that their program order survives: left to itself ptxas re-associates the fully
// unrolled scatter into a GATHER -- it keeps the last K inputs in registers and evaluates every output as one chain of
// K dependent FMAs at the step where it completes, which is latency-bound (one DFMA per ~8 cycles per warp) instead of
// K independent FMAs per step.
__device__ __forceinline__ void fma_acc(double& acc, double w, double v) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(w), "d"(v)); }
__device__ __forceinline__ void fma_acc(float& acc, float w, float v) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(w), "f"(v)); }
__device__ __forceinline__ void mul_acc(double& acc, double w, double v) { asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(acc) : "d"(w), "d"(v)); }
__device__ __forceinline__ void mul_acc(float& acc, float w, float v) { asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(acc) : "f"(w), "f"(v)); }

// Two fp32 values in one 64-bit register, for the packed FFMA2 / FMUL2 of sm_100 (fma.rn.f32x2): one issue slot per two
// FMAs.  The fp32 marches treat two adjacent columns (z marches) or the two 32-column halves of a strip (in-plane window
// march) as one element of this type; taps are broadcast pairs (w, w) and live in uniform registers like the fp64 taps.
struct f32x2 {
    unsigned long long u;
    f32x2() = default;
    __host__ __device__ f32x2(float a, float b) {
        unsigned int x, y;
        memcpy(&x, &a, 4); memcpy(&y, &b, 4);
        u = ((unsigned long long)y << 32) | x;
    }
    __host__ __device__ explicit f32x2(double w) : f32x2((float)w, (float)w) {}
    __host__ __device__ explicit f32x2(int w) : f32x2((float)w, (float)w) {}
    __device__ __forceinline__ float lo() const { return __uint_as_float((unsigned int)u); }
    __device__ __forceinline__ float hi() const { return __uint_as_float((unsigned int)(u >> 32)); }
    __device__ __forceinline__ f32x2 operator-() const { f32x2 r; r.u = u ^ 0x8000000080000000ull; return r; }
};
__device__ __forceinline__ f32x2 operator*(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
    return r;
}
__device__ __forceinline__ void fma_acc(f32x2& acc, f32x2 w, f32x2 v) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.u) : "l"(w.u), "l"(v.u)); }
__device__ __forceinline__ void mul_acc(f32x2& acc, f32x2 w, f32x2 v) { asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(acc.u) : "l"(w.u), "l"(v.u)); }

// One scatter step at static phase PH of an unrolled period P: input v (position p = o + R for the
// output o that completes now) is accumulated into the K outputs it touches.  Returns the completed sum.
// SYM = +1 / -1: the taps are (anti)symmetric, w[K-1-k] = +-w[k] (checked by the host, fast_supported): only the first
// R + 1 of them are read, which halves the registers the taps occupy.
// WARM: the step is one of the first K - 1 of a march (step index == ph): tap k feeds the output that lies ph - k planes
// after the first one, so the taps k > ph only feed outputs before the range -- a triangle of K (K - 1) / 2 useless FMAs
// per march, a third of the work of a 37-tap march over 64 planes -- and are skipped.
template <typename T, int K, int P, int SYM = 0, bool WARM = false>
__device__ __forceinline__ T ring_push(T (&acc)[P], const Taps<T, K>& f, const T v, const int ph) {
    constexpr int R = K / 2;
    const T vn = SYM < 0 ? -v : v;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (WARM && k > ph) continue;
        const int slot = (ph + R - k + 2 * P) % P;   // constant after unrolling
        const bool mirror = SYM != 0 && k > R;
        const T w = f.w[mirror ? K - 1 - k : k];
        if (k == 0) mul_acc(acc[slot], w, v);
        else fma_acc(acc[slot], w, mirror ? vn : v);
    }
    return acc[(ph - R + 2 * P) % P];
}

// ---- the SHIFTING ring -------------------------------------------------------------------------------------------
// acc[i] holds the partial sum of the output that completes i steps from now.  A step with input v is
//     acc[i] = fma(w[K - 1 - i], v, acc[i + 1])   (i = 0 .. K - 2, in this order),     acc[K - 1] = w[0] * v
// -- the FMA writes a different register than it reads, so the ring advances by itself: the same K instructions serve
// every step.  No unrolled period (the statically rotated ring above is P x K FMAs of code: 44 KB for a 49-tap window
// against a 32 KB instruction cache), K accumulators instead of P, any number of steps per stage.  An output receives
// the same operations in the same order as in ring_push (tap 0 by a multiplication, then taps 1 .. K - 1 by FMAs as the
// inputs arrive): the results are bit-identical.
__device__ __forceinline__ void fma_to(double& dst, double w, double v, double src) { asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(dst) : "d"(w), "d"(v), "d"(src)); }
__device__ __forceinline__ void fma_to(float& dst, float w, float v, float src) { asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(dst) : "f"(w), "f"(v), "f"(src)); }
__device__ __forceinline__ void fma_to(f32x2& dst, f32x2 w, f32x2 v, f32x2 src) { asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dst.u) : "l"(w.u), "l"(v.u), "l"(src.u)); }

template <typename T, int K, int SYM>
__device__ __forceinline__ T shift_push(T (&acc)[K], const Taps<T, K>& f, const T v) {
    constexpr int R = K / 2;
    const T vn = SYM < 0 ? -v : v;
#pragma unroll
    for (int i = 0; i < K - 1; ++i) {
        const int k = K - 1 - i;
        const bool mirror = SYM != 0 && k > R;
        fma_to(acc[i], f.w[mirror ? K - 1 - k : k], mirror ? vn : v, acc[i + 1]);
    }
    mul_acc(acc[K - 1], f.w[0], v);
    return acc[0];
}
// The same step at either end of a march: step s lies d = s - nout steps into the tail; tap k feeds the output s - k
// planes after the first one of the range, so only the taps d < k <= s feed outputs inside it (callers may pass the
// largest s and the smallest d of a group of steps: one set of tests for all of them).  The others are skipped in
// groups of G accumulators behind warp-uniform branches: the warm-up triangle at the front, its mirror image at the tail
// and the steps a march is padded by, K (K - 1) / 2 FMAs each.  (A skipped accumulator keeps a stale value: it belongs
// to an output outside the range, which is never stored, and valid sums only ever read valid sums.)
template <typename T, int K, int SYM, int G>
__device__ __forceinline__ T shift_push_edge(T (&acc)[K], const Taps<T, K>& f, const T v, const int s, const int d) {
    constexpr int R = K / 2;
    const T vn = SYM < 0 ? -v : v;
#pragma unroll
    for (int ib = 0; ib < K; ib += G) {
        // accumulators [ib, ib + G) take the taps K - 1 - ib down to K - ib - G
        if (K - 1 - ib > d && K - ib - G <= s) {
#pragma unroll
            for (int i = ib; i < (ib + G < K ? ib + G : K); ++i) {
                const int k = K - 1 - i;
                const bool mirror = SYM != 0 && k > R;
                const T w = f.w[mirror ? K - 1 - k : k];
                if (i == K - 1) mul_acc(acc[i], w, v);
                else fma_to(acc[i], w, mirror ? vn : v, acc[i + 1 < K ? i + 1 : i]);
            }
        }
    }
    return acc[0];
}

// The same step in a stage at the TAIL of a march, with the useless taps skipped in groups of G behind warp-uniform
// branches.  Step s lies d = s - nout steps into the tail; tap k feeds the output s - k planes after the first one of the
// range, so the taps k <= d only feed outputs beyond it: the mirror image of the warm-up triangle, K (K - 1) / 2 FMAs at a
// run-time phase of the unrolled period.  d0 = d of the FIRST step of the stage (one set of tests per stage: a group is
// skipped when none of its taps is useful in any step of the stage).  A skipped group leaves its accumulators stale: they
// belong to outputs outside the range, which are never stored, and a slot is re-initialised by tap 0 of the output that
// claims it next.
template <typename T, int K, int P, int SYM, int G>
__device__ __forceinline__ T ring_push_tail(T (&acc)[P], const Taps<T, K>& f, const T v, const int ph, const int d0) {
    constexpr int R = K / 2;
    const T vn = SYM < 0 ? -v : v;
#pragma unroll
    for (int kb = 0; kb < K; kb += G) {
        if (d0 < kb + G - 1) {                              // kb + G - 1 > d0: the group holds a useful tap in this stage
#pragma unroll
            for (int k = kb; k < (kb + G < K ? kb + G : K); ++k) {
                const int slot = (ph + R - k + 2 * P) % P;
                const bool mirror = SYM != 0 && k > R;
                const T w = f.w[mirror ? K - 1 - k : k];
                if (k == 0) mul_acc(acc[slot], w, v);
                else fma_acc(acc[slot], w, mirror ? vn : v);
            }
        }
    }
    return acc[(ph - R + 2 * P) % P];
}

// ------------------------------------------------------------------------------------------------
// Products + window z pass (calc_flow.py:300-313, z passes): one warp per (channel, y row, 32 x columns, z chunk);
// the product of the channel's two gradient volumes is formed as the values come out of the prefetch ring.
template <typename T, int K>
struct WindowArgs {
    MarchGeom g;
    Taps<T, K> f;
    const T* in[4];   // gradient volumes {dt, dx, dy, dz}
    T* out;           // channel-major output {xx,xy,xz,yy,yz,zz,tx,ty,tz}
    int nch;          // 9
    int edge_skip;    // TMA march: skip the useless taps of the last 2R steps (0: OF3D_NO_TAIL_SKIP, for A/B measurements)
};

// channel -> gradient pair; index into {dt,dx,dy,dz}
__device__ __forceinline__ void channel_pair(int ch, int& a, int& b) {
    const int A[9] = {1, 1, 1, 2, 2, 3, 1, 2, 3};
    const int B[9] = {1, 2, 3, 2, 3, 3, 0, 0, 0};
    a = A[ch]; b = B[ch];
}

template <typename T, int DEPTH, int WPB>
constexpr size_t window_smem() { return (size_t)WPB * Prefetcher<T, 2, DEPTH>::elems_per_warp * sizeof(T); }

template <typename T, int K, int P, int DEPTH, int WPB>
__global__ void __launch_bounds__(WPB * 32) march_window(const WindowArgs<T, K> a) {
    static_assert(P >= K && P % DEPTH == 0, "bad unroll period");
    constexpr int R = K / 2;
    using Pre = Prefetcher<T, 2, DEPTH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const MarchGeom& g = a.g;
    // task = (((other * lane_groups + lg) * n_chunks + chunk) * nch + ch): channels adjacent so that the
    // warps of one block re-use each other's gradient loads in L1
    int64_t task = (int64_t)blockIdx.x * WPB + warp;
    const int64_t ntasks = (int64_t)g.n_other * g.lane_groups * g.n_chunks * a.nch;
    if (task >= ntasks) return;
    const int ch = (int)(task % a.nch); task /= a.nch;
    const int chunk = (int)(task % g.n_chunks); task /= g.n_chunks;
    const int lg = (int)(task % g.lane_groups);
    const int64_t other = task / g.lane_groups;

    const int64_t lane0 = (int64_t)lg * 32;
    const bool lane_ok = lane0 + lane < g.n_lane;
    const int64_t lpos = lane_ok ? lane0 + lane : g.n_lane - 1;   // out-of-range lanes read a valid address
    const int c0 = g.m_begin + chunk * g.chunk;
    const int c1 = min(c0 + g.chunk, g.m_end);
    const int64_t base = other * g.stride_other + lpos;

    Pre pre;
    pre.lbase = reinterpret_cast<T*>(smem_raw) + warp * Pre::elems_per_warp + lane;
    pre.sbase = (uint32_t)__cvta_generic_to_shared(pre.lbase);
    pre.qpos = c0 - R;
    pre.n_march_m1 = (int)g.n_march - 1;
    pre.stride_bytes = g.stride_march * (int64_t)sizeof(T);
    const int64_t off0 = base + clampi(c0 - R, g.n_march) * g.stride_march;
    int ia, ib;
    channel_pair(ch, ia, ib);
    pre.gp[0] = reinterpret_cast<const char*>(a.in[ia] + off0);
    pre.gp[1] = reinterpret_cast<const char*>(a.in[ib] + off0);
    T* outc = a.out + (int64_t)ch * g.vol;

    // The march is padded to whole unroll periods and runs without any branch in the unrolled body: steps
    // before the first complete output (s < 2R) and after the last (s >= nout + 2R) compute as usual on clamped
    // inputs and are masked at the store (one unsigned compare also covers out-of-range lanes).
    const int nout = c1 - c0;
    const unsigned nvalid = lane_ok ? (unsigned)nout : 0u;
    const int nsteps = (nout + 2 * R + P - 1) / P * P;
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) pre.issue(d);

    T acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) acc[i] = T(0);
    // store position of the output completed at step s: c0 + s - 2R (plane 0 of the output is m_begin)
    char* optr = reinterpret_cast<char*>(outc + base) + ((int64_t)(c0 - g.m_begin) - 2 * R) * pre.stride_bytes;

    // The value of step s+1 is read from the ring while step s is being accumulated (software pipelining:
    // the LDS latency hides under the K FMAs of the current step).
    cp_async_wait<DEPTH - 2>();
    T vn0 = pre.read(0, 0), vn1 = pre.read(0, 1);
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const T v = vn0 * vn1;
            pre.issue((ph + DEPTH - 1) % DEPTH);
            cp_async_wait<DEPTH - 2>();            // step s+1 has landed
            vn0 = pre.read((ph + 1) % DEPTH, 0);
            vn1 = pre.read((ph + 1) % DEPTH, 1);
            const T res = ring_push<T, K, P, 1>(acc, a.f, v, ph);
            if ((unsigned)(s0 + ph - 2 * R) < nvalid) *reinterpret_cast<T*>(optr) = res;
            optr += pre.stride_bytes;
        }
    }
    cp_async_wait<0>();
}

}  // namespace of3d
