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
// Marching needs the lanes on the contiguous axis, so filtering along x is done on volumes stored
// (z, x, y) ("T layout", y contiguous).  The kernels that switch layout (first gradient pass, window
// x pass) write through a per-warp shared-memory tile so that both the loads and the stores are full
// coalesced lines.  Pass order: gradients y -> x -> z, window z -> x -> y; separable filters commute,
// so this equals the reference's y -> x -> z up to rounding (1e-16 relative).
#pragma once
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

// One scatter step at static phase PH of an unrolled period P: input v (position p = o + R for the
// output o that completes now) is accumulated into the K outputs it touches.  Returns the completed sum.
template <typename T, int K, int P>
__device__ __forceinline__ T ring_push(T (&acc)[P], const Taps<T, K>& f, const T v, const int ph) {
    constexpr int R = K / 2;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int slot = (ph + R - k + 2 * P) % P;   // constant after unrolling
        if (k == 0) acc[slot] = f.w[0] * v;
        else acc[slot] = fma(f.w[k], v, acc[slot]);
    }
    return acc[(ph - R + 2 * P) % P];
}

// Per-warp transposing tile with one row per phase of the unrolled period (P rows x 33): inside the unrolled
// march a completed output is parked with a single STS at a static address; after each period the tile is
// written out transposed, ONCE, from code that sits outside the unrolled region (keeps the loop inside the
// instruction cache).  Row r of the tile holds output index j0 + r (relative to the chunk start c0); rows
// whose index falls outside [0, nout) are skipped (warm-up steps, tail of the last period).
template <typename T, int P>
__device__ __forceinline__ void tile_flush(const T* __restrict__ t, T* __restrict__ out, int64_t pitch, int64_t c0, int j0,
                                           int nout, int64_t lane0, int64_t n_lane) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    const int ncols = (int)min((int64_t)32, n_lane - lane0);
#pragma unroll
    for (int r0 = 0; r0 < P; r0 += 32) {
        const int r = r0 + lane;
        const int j = j0 + r;
        if (r < P && j >= 0 && j < nout) {
            T* o = out + lane0 * pitch + c0 + j;
            const T* ti = t + r * 33;
#pragma unroll 4
            for (int c = 0; c < ncols; ++c) o[(int64_t)c * pitch] = ti[c];
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Window pass: one stream, K taps.  PROD: the input is the product of two gradient volumes formed on
// the fly (calc_flow.py:300-313).  TR: transposed store (T layout -> N layout).
template <typename T, int K>
struct WindowArgs {
    MarchGeom g;
    Taps<T, K> f;
    const T* in[4];   // PROD: gradient volumes {dt, dx, dy, dz}; else in[0] = base of the channel-major input
    T* out;           // channel-major output
    int nch;          // 9 (3D) or 5 (2D)
    int ndim;
};

// channel -> gradient pair; order {xx,xy,xz,yy,yz,zz,tx,ty,tz} (3D) and {xx,xy,yy,tx,ty} (2D); index into {dt,dx,dy,dz}
__device__ __forceinline__ void channel_pair(int ndim, int ch, int& a, int& b) {
    if (ndim == 3) {
        const int A[9] = {1, 1, 1, 2, 2, 3, 1, 2, 3};
        const int B[9] = {1, 2, 3, 2, 3, 3, 0, 0, 0};
        a = A[ch]; b = B[ch];
    } else {
        const int A[5] = {1, 1, 2, 1, 2};
        const int B[5] = {1, 2, 2, 0, 0};
        a = A[ch]; b = B[ch];
    }
}

// shared-memory bytes of one block
template <typename T, int P, int DEPTH, bool PROD, bool TR, int WPB>
constexpr size_t window_smem() {
    return (size_t)WPB * (Prefetcher<T, PROD ? 2 : 1, DEPTH>::elems_per_warp + (TR ? P * 33 : 0)) * sizeof(T);
}

template <typename T, int K, int P, int DEPTH, bool PROD, bool TR, int WPB>
__global__ void __launch_bounds__(WPB * 32) march_window(const WindowArgs<T, K> a) {
    static_assert(P >= K && P % DEPTH == 0, "bad unroll period");
    constexpr int R = K / 2;
    constexpr int NIN = PROD ? 2 : 1;
    using Pre = Prefetcher<T, NIN, DEPTH>;
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
    const int c0 = chunk * g.chunk;
    const int c1 = min(c0 + g.chunk, (int)g.n_march);
    const int64_t base = other * g.stride_other + lpos;

    T* sm = reinterpret_cast<T*>(smem_raw);
    Pre pre;
    pre.lbase = sm + warp * Pre::elems_per_warp + lane;
    pre.sbase = (uint32_t)__cvta_generic_to_shared(pre.lbase);
    pre.qpos = c0 - R;
    pre.n_march_m1 = (int)g.n_march - 1;
    pre.stride_bytes = g.stride_march * (int64_t)sizeof(T);
    const int64_t off0 = base + clampi(c0 - R, g.n_march) * g.stride_march;
    if (PROD) {
        int ia, ib;
        channel_pair(a.ndim, ch, ia, ib);
        pre.gp[0] = reinterpret_cast<const char*>(a.in[ia] + off0);
        pre.gp[NIN - 1] = reinterpret_cast<const char*>(a.in[ib] + off0);
    } else {
        pre.gp[0] = reinterpret_cast<const char*>(a.in[0] + (int64_t)ch * g.vol + off0);
    }
    T* outc = a.out + (int64_t)ch * g.vol;
    T* tile = sm + WPB * Pre::elems_per_warp + warp * (P * 33) + lane;   // this lane's column of the P-row tile

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
    // plain store position of the output completed at step s: c0 + s - 2R
    char* optr = reinterpret_cast<char*>(outc + base) + ((int64_t)c0 - 2 * R) * pre.stride_bytes;

    // The value of step s+1 is read from the ring while step s is being accumulated (software pipelining:
    // the LDS latency hides under the K FMAs of the current step).
    cp_async_wait<DEPTH - 2>();
    T vn0 = pre.read(0, 0), vn1 = PROD ? pre.read(0, 1) : T(1);
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const T v = PROD ? vn0 * vn1 : vn0;
            pre.issue((ph + DEPTH - 1) % DEPTH);
            cp_async_wait<DEPTH - 2>();            // step s+1 has landed
            vn0 = pre.read((ph + 1) % DEPTH, 0);
            if (PROD) vn1 = pre.read((ph + 1) % DEPTH, 1);
            const T res = ring_push<T, K, P>(acc, a.f, v, ph);
            if (TR) {
                tile[ph * 33] = res;                // output index s - 2R; validity is checked at the flush
            } else {
                if ((unsigned)(s0 + ph - 2 * R) < nvalid) *reinterpret_cast<T*>(optr) = res;
                optr += pre.stride_bytes;
            }
        }
        if (TR) tile_flush<T, P>(tile - lane, outc + other * g.stride_other, g.n_march, c0, s0 - 2 * R, nout, lane0, g.n_lane);
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Gradient passes: several (input, filter) streams marched together.
//   MODE 0  G1   (N layout, march y, TRANSPOSED out): in {dt0, ic}        -> {G dt0, D ic, S ic}
//   MODE 1  G2   (T layout, march x):                 in {A0, A1, A2}     -> {G A0, S A1, D A2, S A2}
//   MODE 2  G3   (T layout, march z):                 in {B0, B1, B2, B3} -> {G B0, S B1, S B2, D B3} = {dt,dy,dx,dz}
//   MODE 3  G2-2D(T layout, march x):                 in {A0, A1, A2}     -> {G A0, S A1, D A2}      = {dt,dy,dx}
// (calc_flow.py:279-288 / 116-122 with the shared sub-results S_y I_c evaluated once.)
template <typename T, int KR, int KS>
struct GradArgs {
    MarchGeom g;
    Taps<T, KR> fG, fD;
    Taps<T, KS> fS;
    const T* in[4];
    T* out[4];
};

template <int MODE> constexpr int grad_nin() { return MODE == 0 ? 2 : (MODE == 2 ? 4 : 3); }
template <int MODE> constexpr int grad_nout() { return (MODE == 0 || MODE == 3) ? 3 : 4; }
template <typename T, int MODE, int P, int DEPTH, int WPB>
constexpr size_t grad_smem() {
    return (size_t)WPB * (Prefetcher<T, grad_nin<MODE>(), DEPTH>::elems_per_warp + (MODE == 0 ? grad_nout<MODE>() * P * 33 : 0)) * sizeof(T);
}

template <typename T, int KR, int KS, int MODE, int P, int DEPTH, int WPB>
__global__ void __launch_bounds__(WPB * 32) march_grad(const GradArgs<T, KR, KS> a) {
    static_assert(P >= KR && KR >= KS && P % DEPTH == 0, "bad unroll period");
    constexpr int R = KR / 2, RS = KS / 2;
    constexpr int NIN = grad_nin<MODE>();
    constexpr int NOUT = grad_nout<MODE>();
    constexpr bool TR = MODE == 0;
    using Pre = Prefetcher<T, NIN, DEPTH>;
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
    const int64_t lpos = lane_ok ? lane0 + lane : g.n_lane - 1;
    const int c0 = chunk * g.chunk;
    const int c1 = min(c0 + g.chunk, (int)g.n_march);
    const int64_t base = other * g.stride_other + lpos;

    // stream st: filter kind (0 = G, 1 = D, both KR taps; 2 = S, KS taps) and input index
    constexpr int kind[4][4] = {{0, 1, 2, 2}, {0, 2, 1, 2}, {0, 2, 2, 1}, {0, 2, 1, 2}};
    constexpr int src[4][4] = {{0, 1, 1, 1}, {0, 1, 2, 2}, {0, 1, 2, 3}, {0, 1, 2, 2}};

    T* sm = reinterpret_cast<T*>(smem_raw);
    Pre pre;
    pre.lbase = sm + warp * Pre::elems_per_warp + lane;
    pre.sbase = (uint32_t)__cvta_generic_to_shared(pre.lbase);
    pre.qpos = c0 - R;
    pre.n_march_m1 = (int)g.n_march - 1;
    pre.stride_bytes = g.stride_march * (int64_t)sizeof(T);
    const int64_t off0 = base + clampi(c0 - R, g.n_march) * g.stride_march;
#pragma unroll
    for (int i = 0; i < NIN; ++i) pre.gp[i] = reinterpret_cast<const char*>(a.in[i] + off0);
    T* tiles = sm + WPB * Pre::elems_per_warp + warp * (NOUT * P * 33);   // NOUT tiles of P rows

    const int nout = c1 - c0;
    const unsigned nvalid = lane_ok ? (unsigned)nout : 0u;
    const int nsteps = (nout + 2 * R + P - 1) / P * P;     // padded to whole periods: no branch in the unrolled body
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) pre.issue(d);

    T acc[NOUT][P];
#pragma unroll
    for (int st = 0; st < NOUT; ++st)
#pragma unroll
        for (int i = 0; i < P; ++i) acc[st][i] = T(0);
    // plain-store byte offsets of the output completed at step s by the wide (c0 + s - 2R) and narrow
    // (c0 + s - R - RS) filters
    int64_t ooffR = base * (int64_t)sizeof(T) + ((int64_t)c0 - 2 * R) * pre.stride_bytes;
    int64_t ooffS = base * (int64_t)sizeof(T) + ((int64_t)c0 - R - RS) * pre.stride_bytes;

    cp_async_wait<DEPTH - 2>();
    T vn[NIN];
#pragma unroll
    for (int i = 0; i < NIN; ++i) vn[i] = pre.read(0, i);
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            T v[NIN];
#pragma unroll
            for (int i = 0; i < NIN; ++i) v[i] = vn[i];
            pre.issue((ph + DEPTH - 1) % DEPTH);
            cp_async_wait<DEPTH - 2>();            // step s+1 has landed: read it while step s is accumulated
#pragma unroll
            for (int i = 0; i < NIN; ++i) vn[i] = pre.read((ph + 1) % DEPTH, i);
            const bool okR = (unsigned)(s0 + ph - 2 * R) < nvalid;
            const bool okS = (unsigned)(s0 + ph - R - RS) < nvalid;
#pragma unroll
            for (int st = 0; st < NOUT; ++st) {
                const int kd = kind[MODE][st];
                const T x = v[src[MODE][st]];
                T res;
                if (kd == 2) res = ring_push<T, KS, P>(acc[st], a.fS, x, ph);
                else res = ring_push<T, KR, P>(acc[st], kd == 0 ? a.fG : a.fD, x, ph);
                if (TR) {
                    tiles[(st * P + ph) * 33 + lane] = res;
                } else if (kd == 2 ? okS : okR) {
                    *reinterpret_cast<T*>(reinterpret_cast<char*>(a.out[st]) + (kd == 2 ? ooffS : ooffR)) = res;
                }
            }
            if (!TR) { ooffR += pre.stride_bytes; ooffS += pre.stride_bytes; }
        }
        if (TR) {
#pragma unroll 1
            for (int st = 0; st < NOUT; ++st)
                tile_flush<T, P>(tiles + st * P * 33, a.out[st] + other * g.stride_other, g.n_march, c0,
                                 s0 - (kind[MODE][st] == 2 ? R + RS : 2 * R), nout, lane0, g.n_lane);
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Last window pass (march y, N layout, lanes on x) fused with the per-voxel solve and reliability
// (calc_flow.py:337-357 / 154-168): the window sums never go back to HBM.
//
// The NCH warps of a block march the NCH channels of the same (z, x-group, y-chunk) and park every
// completed row in shared memory.  After a batch of BR = NR*NCH rows the block synchronises once and each warp
// solves NR rows of the batch, evaluated together so that the long dependent chains of the solve (reciprocal,
// rsqrt, Newton steps of the eigenvalue) overlap (instruction-level parallelism).  The park is double-buffered,
// so one barrier per batch is enough.
template <typename T, int K>
struct SolveArgs {
    MarchGeom g;
    Taps<T, K> f;
    const T* in;   // channel-major, N layout
    T* vx; T* vy; T* vz; T* rel;
};

constexpr int kSolveRowsPerWarp = 3;
template <typename T, int DEPTH, int NCH>
constexpr size_t solve_smem() {
    return (size_t)(2 * kSolveRowsPerWarp * NCH * NCH * 32 + NCH * Prefetcher<T, 1, DEPTH>::elems_per_warp) * sizeof(T);
}

template <typename T, int K, int P, int DEPTH, int NCH>
__global__ void __launch_bounds__(NCH * 32, 1) march_solve(const SolveArgs<T, K> a) {
    constexpr int NR = kSolveRowsPerWarp;
    constexpr int BR = NR * NCH;                            // rows per batch
    static_assert(P >= K && P % BR == 0 && P % DEPTH == 0, "bad unroll period");
    constexpr int R = K / 2;
    using Pre = Prefetcher<T, 1, DEPTH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* park = reinterpret_cast<T*>(smem_raw);               // [2][row in batch][channel][lane]
    const int lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
    const MarchGeom& g = a.g;
    int64_t task = blockIdx.x;
    const int chunk = (int)(task % g.n_chunks); task /= g.n_chunks;
    const int lg = (int)(task % g.lane_groups);
    const int64_t other = task / g.lane_groups;
    const int64_t lane0 = (int64_t)lg * 32;
    const bool lane_ok = lane0 + lane < g.n_lane;
    const int64_t lpos = lane_ok ? lane0 + lane : g.n_lane - 1;
    const int c0 = chunk * g.chunk;
    const int c1 = min(c0 + g.chunk, (int)g.n_march);
    const int64_t base = other * g.stride_other + lpos;
    const int nout = c1 - c0;
    const int nbatch = (nout + BR - 1) / BR;

    Pre pre;
    pre.lbase = park + 2 * BR * NCH * 32 + ch * Pre::elems_per_warp + lane;
    pre.sbase = (uint32_t)__cvta_generic_to_shared(pre.lbase);
    pre.qpos = c0 - R;
    pre.n_march_m1 = (int)g.n_march - 1;
    pre.stride_bytes = g.stride_march * (int64_t)sizeof(T);
    pre.gp[0] = reinterpret_cast<const char*>(a.in + (int64_t)ch * g.vol + base + clampi(c0 - R, g.n_march) * g.stride_march);

    // Outputs complete at steps s >= 2R; output j = s - 2R goes to row j mod BR of batch floor(j / BR).  The
    // march is padded to whole unroll periods and has no branch per step: the steps before the first output fill
    // "batch -1", which is solved like any other and masked at the store, as are the rows past the chunk end.
    const int nsteps = (2 * R + nbatch * BR + P - 1) / P * P;
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) pre.issue(d);

    T acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) acc[i] = T(0);
    constexpr int kRowBias = (2 * R + BR - 1) / BR * BR;    // multiple of BR, >= 2R
    int n = -(kRowBias / BR);                               // batch being filled (negative: warm-up); buffer n & 1
    T* pk = park + ((n & 1) * BR * NCH * 32) + ch * 32 + lane;   // this warp's slot in row 0 of that buffer

    cp_async_wait<DEPTH - 2>();
    T vn = pre.read(0, 0);
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int ph = 0; ph < P; ++ph) {
            const T v = vn;
            pre.issue((ph + DEPTH - 1) % DEPTH);
            cp_async_wait<DEPTH - 2>();            // step s+1 has landed: read it while step s is accumulated
            vn = pre.read((ph + 1) % DEPTH, 0);
            const T res = ring_push<T, K, P>(acc, a.f, v, ph);
            const int row = (ph + kRowBias - 2 * R) % BR;             // == (s - 2R) mod BR since BR | P | s0
            pk[row * NCH * 32] = res;
            if (row == BR - 1) {
                __syncthreads();
                // warp `ch` solves rows ch, ch + NCH, ... of this batch, together
                const T* qb = park + ((n & 1) * BR + ch) * NCH * 32 + lane;
                if (NCH == 9) {
                    Flow3 r[NR];
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const T* qv = qb + i * NCH * NCH * 32;
#ifdef OF3D_EXP_NOSOLVE
                        r[i].vx = qv[0] + qv[32] + qv[64]; r[i].vy = qv[96] + qv[128]; r[i].vz = qv[160] + qv[192]; r[i].rel = qv[224] + qv[256];
#else
                        r[i] = solve3<false>((double)qv[0], (double)qv[32], (double)qv[64], (double)qv[96], (double)qv[128],
                                             (double)qv[160], (double)qv[192], (double)qv[224], (double)qv[256]);
#endif
                    }
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const int jo = n * BR + ch + i * NCH;
                        if (jo >= 0 && jo < nout && lane_ok) {
                            const int64_t idx = base + (int64_t)(c0 + jo) * g.stride_march;
                            a.vx[idx] = (T)r[i].vx; a.vy[idx] = (T)r[i].vy; a.vz[idx] = (T)r[i].vz; a.rel[idx] = (T)r[i].rel;
                        }
                    }
                } else {
                    Flow2 r[NR];
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const T* qv = qb + i * NCH * NCH * 32;
                        r[i] = solve2<false>((double)qv[0], (double)qv[32], (double)qv[64], (double)qv[96], (double)qv[128]);
                    }
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const int jo = n * BR + ch + i * NCH;
                        if (jo >= 0 && jo < nout && lane_ok) {
                            const int64_t idx = base + (int64_t)(c0 + jo) * g.stride_march;
                            a.vx[idx] = (T)r[i].vx; a.vy[idx] = (T)r[i].vy; a.rel[idx] = (T)r[i].rel;
                        }
                    }
                }
                ++n;
                pk = park + ((n & 1) * BR * NCH * 32) + ch * 32 + lane;
            }
        }
    }
    cp_async_wait<0>();
}

}  // namespace of3d
