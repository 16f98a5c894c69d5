// Plane-strip kernels: the two in-plane passes of a separable stage in ONE launch.  Internal header.
//
// A warp owns 32 columns of the contiguous axis x ("c") of one plane and marches along y ("m") in batches of
// RB = 8 rows; everything before a cross-channel step is private to the warp (only __syncwarp):
//   load    cp.async of the warp's RB rows x (32 + Kg - 1) columns into its shared-memory rows (clamp-to-edge on the
//           source address), issued one batch ahead;
//   gather  filter along x: lane (row r = l/4, block b = l%4) produces 8 consecutive outputs of row r from the row in
//           shared memory -- 8 accumulators, 8 + Kg - 1 LDS, 8*Kg FMAs, all addresses and taps static.  Rows are
//           stored "blocked-transposed" (element e at (e % 8) * PITCH + e / 8) so that lanes read consecutive words;
//   march   filter along y: lane = column; the gathered value of each of the RB rows is scattered into K register
//           accumulators (the statically rotated ring of kernels_march.cuh).
//
// strip_conv2         gradient stage, in-plane passes of one gradient volume (calc_flow.py:279-288 / 116-122).
// strip_window_solve  window x pass + y pass + per-voxel solve (calc_flow.py:300-357 / 133-168): the NCH window
//                     sums of a voxel only ever exist in registers and shared memory.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "kernels_march.cuh"
#include "solve.cuh"

namespace of3d {

#ifndef OF3D_EIG32
#define OF3D_EIG32 1
#endif
constexpr bool kEig32 = OF3D_EIG32 != 0;   // fp32 mode: reliability (smallest eigenvalue) in float32 (solve.cuh min_eig_sym3_f32)
constexpr int kStripRB = 8;    // rows per batch
constexpr int kStripXB = 8;    // outputs per lane in the gather phase

template <int K> constexpr int strip_rowlen() { return 32 + K - 1; }
template <int K> constexpr int strip_pitch() { return ((strip_rowlen<K>() + 7) / 8) | 1; }
// row stride = 4 (mod 16) elements: the four rows a half-warp reads in the gather fall on disjoint banks
template <int K> constexpr int strip_rowstride() { return (8 * strip_pitch<K>() + 11) / 16 * 16 + 4; }
constexpr int kXwRow = 32 + 4;                                               // one pad element every 8 columns

// Geometry shared by both kernels: N layout (other = z, march = y, contiguous = x)
struct StripGeom {
    int64_t vol;            // elements per volume
    int64_t stride_m;       // element stride of y
    int64_t stride_o;       // element stride of z
    int n_c, n_m, n_o;      // extents x, y, z
    int chunk, n_chunks;    // outputs per task along y
};

// Warp-private staging: one or two row buffers (blocked-transposed) and the gathered rows
// (NBUF = 2: two batches of rows in flight -- the HBM-bound gradient stage needs the deeper prefetch)
template <typename T, int KG, int NIN, int NBUF = 1>
struct StripStage {
    static constexpr int ROWLEN = strip_rowlen<KG>(), PITCH = strip_pitch<KG>(), ROWSTRIDE = strip_rowstride<KG>();
    static constexpr int NLOAD = (ROWLEN + 7) / 8;
    static constexpr int BUF = NIN * kStripRB * ROWSTRIDE;
    static constexpr int elems = NBUF * BUF + kStripRB * kXwRow;
    T* rows;                // [NBUF][NIN][RB][ROWSTRIDE]; batch b lives in buffer b % NBUF
    T* xw;                  // [RB][kXwRow]
    uint32_t rows_s;
    const T* src[NIN];      // plane base of each input (already offset to z and channel)
    int c0;                 // first column this lane loads (cw0 - R + lane % 8)
    int ncm1, nmm1, mfirst; // clamp limits; row index of batch 0, row 0
    int64_t stride_m;

    __device__ __forceinline__ void init(T* smem_warp, int cw0, int m_first, const StripGeom& g) {
        rows = smem_warp;
        xw = smem_warp + NBUF * BUF;
        rows_s = (uint32_t)__cvta_generic_to_shared(rows);
        const int lane = threadIdx.x & 31;
        c0 = cw0 - KG / 2 + (lane & 7);
        ncm1 = g.n_c - 1; nmm1 = g.n_m - 1; mfirst = m_first; stride_m = g.stride_m;
    }
    // lane (l8 = lane % 8, rr = lane / 8) fetches elements e = l8 + 8 i of rows rr and rr + 4; element e lives at
    // (e % 8) * PITCH + e / 8 = l8 * PITCH + i
    __device__ __forceinline__ void issue(int b) {
        const int lane = threadIdx.x & 31, l8 = lane & 7, lrr = lane >> 3;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int r = lrr + 4 * p;
            const int m = max(0, min(mfirst + b * kStripRB + r, nmm1));
            const uint32_t dst = rows_s + (uint32_t)(((b % NBUF) * BUF + r * ROWSTRIDE + l8 * PITCH) * sizeof(T));
#pragma unroll
            for (int i = 0; i < NLOAD; ++i) {
                if (ROWLEN % 8 == 0 || l8 + 8 * i < ROWLEN) {
                    const int c = max(0, min(c0 + 8 * i, ncm1));
#pragma unroll
                    for (int q = 0; q < NIN; ++q)
                        cp_async_elem<T>(dst + (uint32_t)((q * kStripRB * ROWSTRIDE + i) * sizeof(T)), src[q] + (int64_t)m * stride_m + c);
                }
            }
        }
        cp_async_commit();
    }
    // gather with filter f (KF <= KG taps, centred in the loaded halo) of input q (or the product of inputs 0 and 1),
    // result to xw
    template <int KF, bool PROD>
    __device__ __forceinline__ void gather(const Taps<T, KF>& f, int q, int b = 0) {
        constexpr int XB = kStripXB, OFF = (KG - KF) / 2;
        const int lane = threadIdx.x & 31, g_r = lane >> 2, g_b = lane & 3;
        const T* s0 = rows + (b % NBUF) * BUF + (q * kStripRB + g_r) * ROWSTRIDE + g_b;
        T ga[XB];
#pragma unroll
        for (int i = 0; i < XB; ++i) ga[i] = T(0);
#pragma unroll
        for (int mm = OFF; mm < XB + KF - 1 + OFF; ++mm) {
            T v = s0[(mm & 7) * PITCH + (mm >> 3)];
            if (PROD) v *= s0[kStripRB * ROWSTRIDE + (mm & 7) * PITCH + (mm >> 3)];
#pragma unroll
            for (int i = 0; i < XB; ++i) {
                const int k = mm - OFF - i;
                if (k >= 0 && k < KF) ga[i] = fma(f.w[k], v, ga[i]);
            }
        }
        T* d = xw + g_r * kXwRow + g_b * (XB + 1);
#pragma unroll
        for (int i = 0; i < XB; ++i) d[i] = ga[i];
    }
    __device__ __forceinline__ T gathered(int r) const {
        const int lane = threadIdx.x & 31;
        return xw[r * kXwRow + lane + lane / 8];
    }
};

// Row-major variant of the staging for 8-byte elements (fp64): rows are fetched with 16-byte cp.async (two elements per
// lane and instruction, half the LDGSTS of the element-wise stage -- the gradient strips ran at 72-78 % L1/LSU
// utilisation) and gathered with 128-bit shared loads.  Lane (row l & 7, block l >> 3) makes the eight lanes of a
// quarter-warp read eight different rows; a row stride = 2 (mod 4) elements puts them on eight different 16-byte bank
// groups.  Chunks lie on the even-column grid (the volume's row length must be even and its base 16-byte aligned), so a
// chunk is either entirely inside the row or entirely outside: outside chunks are fetched from the clamped chunk and the
// edge strips then overwrite their out-of-range columns with the edge value (scipy's mode='nearest').
template <typename T, int KG>
struct StripStage16 {
    static_assert(sizeof(T) == 8, "16-byte staging is written for 8-byte elements");
    static constexpr int R = KG / 2, AL = R & 1;                   // AL: extra leading column -> even first column
    static constexpr int ROWLEN = (32 + KG - 1 + AL + 1) / 2 * 2;  // elements fetched per row
    static constexpr int NCHUNK = ROWLEN / 2;
    static constexpr int ROWSTRIDE = ROWLEN + ((ROWLEN % 4 == 2) ? 0 : 2);
    static constexpr int BUF = kStripRB * ROWSTRIDE;
    static constexpr int elems = 2 * BUF + kStripRB * kXwRow;       // two batches in flight + the gathered rows
    T* rows;                // [2][RB][ROWSTRIDE]; batch b lives in buffer b & 1
    T* xw;                  // [RB][kXwRow]
    uint32_t rows_s;
    const T* src;           // plane base of the input
    int col0;               // first (even) column of the staged rows: cw0 - R - AL
    int n_c, nmm1, mfirst;
    int64_t stride_m;

    __device__ __forceinline__ void init(T* smem_warp, int cw0, int m_first, const StripGeom& g) {
        rows = smem_warp;
        xw = smem_warp + 2 * BUF;
        rows_s = (uint32_t)__cvta_generic_to_shared(rows);
        col0 = cw0 - R - AL;
        n_c = g.n_c; nmm1 = g.n_m - 1; mfirst = m_first; stride_m = g.stride_m;
    }
    // lane (row = lane / 4, q = lane % 4) fetches chunks q, q + 4, ... of its row
    __device__ __forceinline__ void issue(int b) {
        const int lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
        const int m = max(0, min(mfirst + b * kStripRB + r, nmm1));
        const T* row = src + (int64_t)m * stride_m;
        const uint32_t dst = rows_s + (uint32_t)(((b & 1) * BUF + r * ROWSTRIDE + 2 * q) * sizeof(T));
#pragma unroll
        for (int i = 0; i < (NCHUNK + 3) / 4; ++i) {
            if (NCHUNK % 4 == 0 || q + 4 * i < NCHUNK) {
                const int c = max(0, min(col0 + 2 * (q + 4 * i), n_c - 2));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(8 * i * sizeof(T))), "l"(row + c) : "memory");
            }
        }
        cp_async_commit();
    }
    // edge strips: columns outside [0, n_c) take the edge value
    __device__ __forceinline__ void fix_edges(int b) {
        if (col0 >= 0 && col0 + ROWLEN <= n_c) return;
        const int lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
        T* row = rows + (b & 1) * BUF + r * ROWSTRIDE;
        if (col0 < 0) {
            const T v = row[-col0];
            for (int e = q; e < -col0; e += 4) row[e] = v;
        }
        const int last = n_c - 1 - col0;                            // element index of the last valid column
        if (last < ROWLEN - 1) {
            const T v = row[max(last, 0)];
            for (int e = last + 1 + q; e < ROWLEN; e += 4) row[e] = v;
        }
        __syncwarp();
    }
    // gather with filter f (KF <= KG taps, centred in the staged halo), result to xw
    // (SYM = +1 / -1: only the first half of the (anti)symmetric taps is read, cf. ring_push)
    template <int KF, int SYM>
    __device__ __forceinline__ void gather(const Taps<T, KF>& f, int b) {
        constexpr int XB = kStripXB, OFF = AL + (KG - KF) / 2;      // element of tap 0 of output 0 of block 0
        constexpr int EB = OFF & ~1, SKIP = OFF & 1;                // pairs start at the even element EB
        constexpr int NPAIR = (SKIP + XB + KF - 1 + 1) / 2;
        const int lane = threadIdx.x & 31, g_r = lane & 7, g_b = lane >> 3;
        const double2* s0 = reinterpret_cast<const double2*>(rows + (b & 1) * BUF + g_r * ROWSTRIDE + EB + g_b * XB);
        T ga[XB];
#pragma unroll
        for (int i = 0; i < XB; ++i) ga[i] = T(0);
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) {
            const double2 v2 = s0[p];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int mm = 2 * p + h - SKIP;                    // input element relative to tap 0 of output 0
                const T v = h ? v2.y : v2.x;
                const T vn = SYM < 0 ? -v : v;
#pragma unroll
                for (int i = 0; i < XB; ++i) {
                    const int k = mm - i;
                    if (mm >= 0 && k >= 0 && k < KF) {
                        const bool mirror = k > KF / 2;
                        ga[i] = fma(f.w[mirror ? KF - 1 - k : k], mirror ? vn : v, ga[i]);
                    }
                }
            }
        }
        T* d = xw + g_r * kXwRow + g_b * (XB + 1);
#pragma unroll
        for (int i = 0; i < XB; ++i) d[i] = ga[i];
    }
    __device__ __forceinline__ T gathered(int r) const {
        const int lane = threadIdx.x & 31;
        return xw[r * kXwRow + lane + lane / 8];
    }
};

// ------------------------------------------------------------------------------------------------
// In-plane passes of one gradient volume: gather with the x filter fg (KG taps), march with the y filter fa (KA taps).
// The gradient stage is four instantiations (calc_flow.py:279-288 x and y passes; z ran first, kernels_tz.cuh):
//     dt = G_y G_x C      dy = D_y S_x A      dx = S_y D_x A      dz = S_y S_x Bz
// and in 2D (calc_flow.py:116-122)  dt = G_y G_x dt0, dy = D_y S_x Ic, dx = S_y D_x Ic.  Every instantiation stages
// only the halo its own x filter needs and unrolls only the period its own y filter needs.
template <typename T, int KG, int KA>
struct Conv2Args {
    StripGeom g;
    Taps<T, KG> fg;
    Taps<T, KA> fa;
    const T* in;
    T* out;
};

template <typename T, int KG, bool V16 = false>
constexpr size_t conv2_smem(int wpb) {
    if constexpr (V16) return (size_t)wpb * StripStage16<T, KG>::elems * sizeof(T);
    else return (size_t)wpb * StripStage<T, KG, 1, 2>::elems * sizeof(T);
}

// SYMG / SYMA: +1 symmetric, -1 antisymmetric taps (only half of them are read)
// SHIFT: y march on the shifting ring (kernels_march.cuh shift_push): KA accumulators, one copy of the batch; MINB: blocks per SM
template <typename T, int KG, int SYMG, int KA, int SYMA, int WPB, bool V16 = false, bool SHIFT = false, int MINB = 1>
__global__ void __launch_bounds__(WPB * 32, MINB) strip_conv2(const Conv2Args<T, KG, KA> a) {
    constexpr int RB = kStripRB, RA = KA / 2, P = (KA + RB - 1) / RB * RB;
    using Stage = typename std::conditional<V16, StripStage16<typename std::conditional<V16, T, double>::type, KG>, StripStage<T, KG, 1, 2>>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const StripGeom& g = a.g;
    int64_t task = (int64_t)blockIdx.x * WPB + warp;
    const int nstrips = (g.n_c + 31) / 32;
    if (task >= (int64_t)nstrips * g.n_chunks * g.n_o) return;
    const int strip = (int)(task % nstrips); task /= nstrips;
    const int chunk = (int)(task % g.n_chunks);
    const int o = (int)(task / g.n_chunks);
    const int cw0 = strip * 32;
    const int m0 = chunk * g.chunk;
    const int nout = min(m0 + g.chunk, g.n_m) - m0;
    const int nsteps = (nout + 2 * RA + P - 1) / P * P;
    const unsigned nvalid = (cw0 + lane < g.n_c) ? (unsigned)nout : 0u;

    Stage st;
    st.init(reinterpret_cast<T*>(smem_raw) + warp * Stage::elems, cw0, m0 - RA, g);
    if constexpr (V16) st.src = a.in + (int64_t)o * g.stride_o;
    else st.src[0] = a.in + (int64_t)o * g.stride_o;
    T* const op = a.out + (int64_t)o * g.stride_o + cw0 + lane;

    st.issue(0);
    st.issue(1);
    if constexpr (SHIFT) {
        T accs[KA];
#pragma unroll
        for (int i = 0; i < KA; ++i) accs[i] = T(0);
        const int nb = (nout + 2 * RA + RB - 1) / RB;
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            cp_async_wait<1>();
            __syncwarp();
            if constexpr (V16) {
                st.fix_edges(b);
                st.template gather<KG, SYMG>(a.fg, b);
            } else {
                st.template gather<KG, false>(a.fg, 0, b);
            }
            __syncwarp();
            st.issue(b + 2);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int s = b * RB + r;
                const T res = shift_push<T, KA, SYMA>(accs, a.fa, st.gathered(r));
                if ((unsigned)(s - 2 * RA) < nvalid) op[(int64_t)(m0 + s - 2 * RA) * g.stride_m] = res;
            }
            __syncwarp();
        }
        cp_async_wait<0>();
        return;
    }
    T acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) acc[i] = T(0);
    int b = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int bi = 0; bi < P / RB; ++bi, ++b) {
            cp_async_wait<1>();                                               // batch b has landed, b + 1 is in flight
            __syncwarp();
            if constexpr (V16) {
                st.fix_edges(b);
                st.template gather<KG, SYMG>(a.fg, b);
            } else {
                st.template gather<KG, false>(a.fg, 0, b);
            }
            __syncwarp();
            st.issue(b + 2);                                                  // into the buffer just consumed
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int ph = bi * RB + r;
                const int s = s0 + ph;
                const T res = ring_push<T, KA, P, SYMA>(acc, a.fa, st.gathered(r), ph);
                if ((unsigned)(s - 2 * RA) < nvalid) op[(int64_t)(m0 + s - 2 * RA) * g.stride_m] = res;
            }
            __syncwarp();                                                     // gathered rows consumed before the next gather
        }
    }
    cp_async_wait<0>();
}

// Two gradients from ONE staged input (3D: dy = D_y S_x A and dx = S_y D_x A share A; 2D: the same on Ic): the rows are
// staged once with the halo of the wider x filter, gathered twice and marched twice.  Saves one read of the input volume
// and one launch against two strip_conv2 launches.  outA = march fa1 (KR taps, antisymmetric D) of gather fg1 (KS taps, S);
// outB = march fa2 (KS taps, S) of gather fg2 (KR taps, antisymmetric D).
template <typename T, int KR, int KS>
struct Conv2DualArgs {
    StripGeom g;
    Taps<T, KR> fD;
    Taps<T, KS> fS;
    const T* in;
    T* out_dy;      // D_y S_x in
    T* out_dx;      // S_y D_x in
};

template <typename T, int KR, int KS, int WPB, bool V16 = false, bool SHIFT = false, int MINB = 1>
__global__ void __launch_bounds__(WPB * 32, MINB) strip_conv2_dual(const Conv2DualArgs<T, KR, KS> a) {
    constexpr int RB = kStripRB, R = KR / 2, RS = KS / 2, P = (KR + RB - 1) / RB * RB;
    using Stage = typename std::conditional<V16, StripStage16<typename std::conditional<V16, T, double>::type, KR>, StripStage<T, KR, 1, 2>>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const StripGeom& g = a.g;
    int64_t task = (int64_t)blockIdx.x * WPB + warp;
    const int nstrips = (g.n_c + 31) / 32;
    if (task >= (int64_t)nstrips * g.n_chunks * g.n_o) return;
    const int strip = (int)(task % nstrips); task /= nstrips;
    const int chunk = (int)(task % g.n_chunks);
    const int o = (int)(task / g.n_chunks);
    const int cw0 = strip * 32;
    const int m0 = chunk * g.chunk;
    const int nout = min(m0 + g.chunk, g.n_m) - m0;
    const int nsteps = (nout + 2 * R + P - 1) / P * P;
    const unsigned nvalid = (cw0 + lane < g.n_c) ? (unsigned)nout : 0u;

    Stage st;
    st.init(reinterpret_cast<T*>(smem_raw) + warp * Stage::elems, cw0, m0 - R, g);
    if constexpr (V16) st.src = a.in + (int64_t)o * g.stride_o;
    else st.src[0] = a.in + (int64_t)o * g.stride_o;
    const int64_t obase = (int64_t)o * g.stride_o + cw0 + lane;
    T* const oy = a.out_dy + obase;
    T* const ox = a.out_dx + obase;

    st.issue(0);
    st.issue(1);
    if constexpr (SHIFT) {
        T accYs[KR], accXs[KS];                                               // 26 accumulators for 19 / 7 taps instead of 48
#pragma unroll
        for (int i = 0; i < KR; ++i) accYs[i] = T(0);
#pragma unroll
        for (int i = 0; i < KS; ++i) accXs[i] = T(0);
        const int nb = (nout + 2 * R + RB - 1) / RB;
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            cp_async_wait<1>();
            __syncwarp();
            if constexpr (V16) { st.fix_edges(b); st.template gather<KS, 1>(a.fS, b); }
            else st.template gather<KS, false>(a.fS, 0, b);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < RB; ++r) {                                     // dy = D_y (S_x in): row m0 + s - 2R
                const int s = b * RB + r;
                const T res = shift_push<T, KR, -1>(accYs, a.fD, st.gathered(r));
                if ((unsigned)(s - 2 * R) < nvalid) oy[(int64_t)(m0 + s - 2 * R) * g.stride_m] = res;
            }
            __syncwarp();
            if constexpr (V16) st.template gather<KR, -1>(a.fD, b);
            else st.template gather<KR, false>(a.fD, 0, b);
            __syncwarp();
            st.issue(b + 2);
#pragma unroll
            for (int r = 0; r < RB; ++r) {                                     // dx = S_y (D_x in): row m0 + s - R - RS
                const int s = b * RB + r;
                const T res = shift_push<T, KS, 1>(accXs, a.fS, st.gathered(r));
                if ((unsigned)(s - R - RS) < nvalid) ox[(int64_t)(m0 + s - R - RS) * g.stride_m] = res;
            }
            __syncwarp();
        }
        cp_async_wait<0>();
        return;
    }
    T accY[P], accX[P];
#pragma unroll
    for (int i = 0; i < P; ++i) { accY[i] = T(0); accX[i] = T(0); }
    int b = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int bi = 0; bi < P / RB; ++bi, ++b) {
            cp_async_wait<1>();                                               // batch b has landed, b + 1 is in flight
            __syncwarp();
            if constexpr (V16) { st.fix_edges(b); st.template gather<KS, 1>(a.fS, b); }
            else st.template gather<KS, false>(a.fS, 0, b);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < RB; ++r) {                                     // dy = D_y (S_x in): row m0 + s - 2R
                const int s = s0 + bi * RB + r;
                const T res = ring_push<T, KR, P, -1>(accY, a.fD, st.gathered(r), bi * RB + r);
                if ((unsigned)(s - 2 * R) < nvalid) oy[(int64_t)(m0 + s - 2 * R) * g.stride_m] = res;
            }
            __syncwarp();                                                     // gathered rows consumed
            if constexpr (V16) st.template gather<KR, -1>(a.fD, b);
            else st.template gather<KR, false>(a.fD, 0, b);
            __syncwarp();
            st.issue(b + 2);                                                  // into the buffer both gathers have consumed
#pragma unroll
            for (int r = 0; r < RB; ++r) {                                     // dx = S_y (D_x in): row m0 + s - R - RS
                const int s = s0 + bi * RB + r;
                const T res = ring_push<T, KS, P, 1>(accX, a.fS, st.gathered(r), bi * RB + r);
                if ((unsigned)(s - R - RS) < nvalid) ox[(int64_t)(m0 + s - R - RS) * g.stride_m] = res;
            }
            __syncwarp();
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Window x pass + y pass + solve.  A block has NCH * NHALF warps: warp (channel ch, half h) owns columns
// [cs0 + 32 h, +32) of channel ch.  After gather and march a completed row is parked in shared memory; ONE block
// barrier per batch, then the voxels of the RB x TX batch are solved (calc_flow.py:337-357 / 154-168) in units of 32
// voxels handed to the warps so that the FP64 work per scheduler (warp % 4) is even.  The park is double-buffered:
// warps may run a batch apart, so the latency-bound solve of one warp overlaps the FMA-bound gather/march of others.
// PROD (2D): the inputs are the gradient volumes and the products are formed in the gather (calc_flow.py:133-141).
template <typename T, int K>
struct StripArgs {
    StripGeom g;
    Taps<T, K> f;
    const T* in[4];         // !PROD: in[0] = channel-major window sums of the z pass; PROD: gradients {dt, dx, dy, -}
    T* vx; T* vy; T* vz; T* rel;
    int rel_f32;            // fp64 only: `rel` is a float buffer, the dtype the reference returns in 3D (calc_flow.py:355-357)
};

template <int NCH, int NHALF> constexpr int strip_parkrow() { return NCH * NHALF * 32 + 2; }
template <typename T, int K, int NCH, int NHALF, bool PROD>
constexpr size_t strip_smem() {
    return (size_t)(NCH * NHALF * StripStage<T, K, PROD ? 2 : 1>::elems + 2 * kStripRB * strip_parkrow<NCH, NHALF>()) * sizeof(T);
}

// ui-th solve unit (32 voxels) of warp `warp` or -1; recomputed at every use (a table indexed by ui lives in local
// memory -- its loads were 5 % of the kernel's stall samples -- and four registers per thread are not to spare)
template <int NCH, int NHALF>
__device__ __forceinline__ int solve_unit(int warp, int ui) {
    const int q = warp >> 2, sm = warp & 3;            // q-th warp of scheduler sm
    if (NCH == 9 && NHALF == 2) {                      // 18 warps (5,5,4,4), 16 units: 3,3,5,5 (one unit per warp --
                                                       // 4,4,4,4 -- measured 15 % slower, 2,2,6,6 4 % slower)
        if (sm >= 2) return ui == 0 ? (sm - 2) * 5 + q : ((ui == 1 && q == 0) ? (sm - 2) * 5 + 4 : -1);
        return (ui == 0 && q < 3) ? 10 + sm * 3 + q : -1;
    } else if (NCH == 9 && NHALF == 1) {               // 9 warps (3,2,2,2), 8 units: 0,3,3,2
        if (sm == 1) return ui == 0 ? q : ((ui == 1 && q == 0) ? 2 : -1);
        if (sm == 2) return ui == 0 ? 3 + q : ((ui == 1 && q == 0) ? 5 : -1);
        if (sm == 3) return ui == 0 ? 6 + q : -1;
        return -1;
    } else if (NCH == 5 && NHALF == 2) {               // 10 warps (3,3,2,2), 16 units: 0,0,8,8
        return (sm >= 2 && ui < 4) ? (sm - 2) * 8 + q * 4 + ui : -1;
    } else {                                           // generic: round robin
        const int u = warp + ui * NCH * NHALF;
        return u < kStripRB * NHALF ? u : -1;
    }
}

// Solve the 32 voxels of unit `unit` of batch b (row v / TX, column v % TX of the parked batch) and store the results.
// Solve the voxels of NU units of batch b at once (row v / TX, column v % TX of the parked batch; NU independent
// dependency chains per lane) and store the results.
template <typename T, int NCH, int TX, int R, int NU = 1>
__device__ __forceinline__ void solve_unit_voxels(const T* park, int parkrow, int b, int unit, int lane, int cs0, int nout, int n_c,
                                                  int64_t plane_off, int64_t stride_m, int m0, T* vx, T* vy, T* vz, T* rel, int rel_f32) {
    constexpr int RB = kStripRB;
    bool ok[NU];
    const T* qv[NU];
    int64_t idx[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int v = (unit + u) * 32 + lane;
        const int s_i = v / TX, s_col = v % TX;
        const int s_c = cs0 + s_col;
        const int j = b * RB - 2 * R + s_i;
        ok[u] = j >= 0 && j < nout && s_c < n_c;
        const int prow = ((j % RB) + RB) % RB;
        qv[u] = park + ((b & 1) * RB + prow) * parkrow + s_col;
        idx[u] = plane_off + (int64_t)(m0 + j) * stride_m + s_c;
    }
    if (NCH == 9) {
        Flow3 rr[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u)           // parked values of masked voxels are finite garbage at worst: solved, not stored
            rr[u] = solve3<false, sizeof(T) == 4 && kEig32>((double)qv[u][0], (double)qv[u][TX], (double)qv[u][2 * TX], (double)qv[u][3 * TX], (double)qv[u][4 * TX],
                                  (double)qv[u][5 * TX], (double)qv[u][6 * TX], (double)qv[u][7 * TX], (double)qv[u][8 * TX]);
#pragma unroll
        for (int u = 0; u < NU; ++u)
            if (ok[u]) {
                vx[idx[u]] = (T)rr[u].vx; vy[idx[u]] = (T)rr[u].vy; vz[idx[u]] = (T)rr[u].vz;
                if (sizeof(T) == 8 && rel_f32) reinterpret_cast<float*>(rel)[idx[u]] = (float)rr[u].rel;   // rounded once
                else rel[idx[u]] = (T)rr[u].rel;
            }
    } else {
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            if (!ok[u]) continue;
            const Flow2 rr = solve2<false>((double)qv[u][0], (double)qv[u][TX], (double)qv[u][2 * TX], (double)qv[u][3 * TX], (double)qv[u][4 * TX]);
            vx[idx[u]] = (T)rr.vx; vy[idx[u]] = (T)rr.vy;
            if (sizeof(T) == 8 && rel_f32) reinterpret_cast<float*>(rel)[idx[u]] = (float)rr.rel;
            else rel[idx[u]] = (T)rr.rel;
        }
    }
}

// y march of one batch (rows BI * RB .. BI * RB + RB - 1 of the unroll period): scatter the gathered rows into the ring
// and park the completed rows (output j = s - 2R lives in park row j mod RB)
template <typename T, int K, int P, int BI, typename Stage>
__device__ __forceinline__ void strip_march_batch(T (&acc)[P], const Taps<T, K>& f, const Stage& st, T* pk, int parkrow) {
    constexpr int RB = kStripRB, R = K / 2;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const T res = ring_push<T, K, P, 1>(acc, f, st.gathered(r), BI * RB + r);
        constexpr int kBias = (2 * R + RB - 1) / RB * RB;
        pk[((r + kBias - 2 * R) % RB) * parkrow] = res;
    }
}

// COMPACT (fp64 windows longer than 25 taps): one copy of the gather and of the solve in the instruction stream, the
// batch of the period selected by a uniform switch -- the fully unrolled period of the 49-tap kernel is 155 KB of SASS
// (7 copies of a 392-DFMA gather) and stalls 1.6 cycles per issue on instruction fetch.
// SHIFT: the y march runs on the shifting accumulator ring (kernels_march.cuh shift_push): K accumulators instead of P,
// one copy of gather + march + solve in the instruction stream whatever the window length.
template <typename T, int K, int P, int NCH, int NHALF, bool PROD, bool SHIFT = false>
__global__ void __launch_bounds__(NCH * NHALF * 32, 1) strip_window_solve(const StripArgs<T, K> a) {
    constexpr int RB = kStripRB, R = K / 2, TX = 32 * NHALF;
    constexpr int PARKROW = strip_parkrow<NCH, NHALF>();
    static_assert(P >= K && P % RB == 0, "unroll period must cover the taps and be a multiple of the batch");
    static_assert(NCH * NHALF >= 4 || true, "");
    using Stage = StripStage<T, K, PROD ? 2 : 1>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = warp % NCH, half = warp / NCH;
    T* park = smem + NCH * NHALF * Stage::elems;                             // [2][RB][PARKROW], shared by the block
    const StripGeom& g = a.g;

    int task = blockIdx.x;
    const int nstrips = (g.n_c + TX - 1) / TX;
    const int strip = task % nstrips; task /= nstrips;
    const int chunk = task % g.n_chunks;
    const int o = task / g.n_chunks;
    const int cs0 = strip * TX;                                              // first column of the block's strip
    const int m0 = chunk * g.chunk;
    const int nout = min(m0 + g.chunk, g.n_m) - m0;
    const int nsteps = (nout + 2 * R + P - 1) / P * P;                       // whole unroll periods

    Stage st;
    st.init(smem + warp * Stage::elems, cs0 + 32 * half, m0 - R, g);
    if (PROD) {
        // channel -> gradient pair {xx,xy,yy,tx,ty}; a.in = {dt, dx, dy}
        const int ia = (ch == 2 || ch == 4) ? 2 : 1;
        const int ib = ch == 0 ? 1 : (ch <= 2 ? 2 : 0);
        st.src[0] = a.in[ia] + (int64_t)o * g.stride_o;
        st.src[PROD ? 1 : 0] = a.in[ib] + (int64_t)o * g.stride_o;
    } else {
        st.src[0] = a.in[0] + (int64_t)o * g.stride_o + (int64_t)ch * g.vol;
    }
    T* const m_dst = park + ch * TX + 32 * half + lane;

    auto solve_batch = [&](int bb) {
#pragma unroll 1
        for (int ui = 0; ui < 4; ++ui) {
            const int unit = solve_unit<NCH, NHALF>(warp, ui);
            if (unit < 0) break;
            solve_unit_voxels<T, NCH, TX, R>(park, PARKROW, bb, unit, lane, cs0, nout, g.n_c, (int64_t)o * g.stride_o, g.stride_m, m0,
                                             a.vx, a.vy, a.vz, a.rel, a.rel_f32);
        }
    };

    st.issue(0);
    if constexpr (SHIFT) {
        T accs[K];
#pragma unroll
        for (int i = 0; i < K; ++i) accs[i] = T(0);
        const int nb = (nout + 2 * R + RB - 1) / RB;                         // whole batches
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            cp_async_wait<0>();
            __syncwarp();                                                     // this warp's rows of batch b have landed
            st.template gather<K, PROD>(a.f, 0);
            __syncwarp();                                                     // gathered rows visible; input rows free
            st.issue(b + 1);                                                  // prefetch (clamped addresses: always valid)
            T* pk = m_dst + (b & 1) * RB * PARKROW;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const T res = shift_push<T, K, 1>(accs, a.f, st.gathered(r));
                constexpr int kBias = (2 * R + RB - 1) / RB * RB;
                pk[((r + kBias - 2 * R) % RB) * PARKROW] = res;
            }
            __syncthreads();                                                  // one batch of outputs parked by all channels
            solve_batch(b);
        }
        cp_async_wait<0>();
        return;
    }
    T acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) acc[i] = T(0);
    int b = 0;
    constexpr bool COMPACT = sizeof(T) == 8 && K > 25;
    if constexpr (COMPACT) {
        static_assert(P / RB <= 7, "extend the switch");
        const int nb = nsteps / RB;
#pragma unroll 1
        for (; b < nb; ++b) {
            cp_async_wait<0>();
            __syncwarp();
            st.template gather<K, PROD>(a.f, 0);
            __syncwarp();
            st.issue(b + 1);
            T* pk = m_dst + (b & 1) * RB * PARKROW;
            switch (b % (P / RB)) {
#define OF3D_CASE(i) case i: if constexpr (i < P / RB) strip_march_batch<T, K, P, (i < P / RB ? i : 0)>(acc, a.f, st, pk, PARKROW); break;
                OF3D_CASE(0) OF3D_CASE(1) OF3D_CASE(2) OF3D_CASE(3) OF3D_CASE(4) OF3D_CASE(5) OF3D_CASE(6)
#undef OF3D_CASE
                default: break;
            }
            __syncthreads();
            solve_batch(b);
        }
        cp_async_wait<0>();
        return;
    }
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int bi = 0; bi < P / RB; ++bi, ++b) {
            cp_async_wait<0>();
            __syncwarp();                                                     // this warp's rows of batch b have landed
            st.template gather<K, PROD>(a.f, 0);
            __syncwarp();                                                     // gathered rows visible; input rows free
            st.issue(b + 1);                                                  // prefetch (clamped addresses: always valid)
            // ---------------- march along y: output j = s - 2R lives in park row j mod RB of buffer b & 1
            {
                T* pk = m_dst + (b & 1) * RB * PARKROW;
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const T res = ring_push<T, K, P, 1>(acc, a.f, st.gathered(r), bi * RB + r);
                    constexpr int kBias = (2 * R + RB - 1) / RB * RB;
                    pk[((r + kBias - 2 * R) % RB) * PARKROW] = res;
                }
            }
            __syncthreads();                                                  // one batch of outputs parked by all channels
            solve_batch(b);                                                   // the RB x TX outputs of batch b
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// 2D: products + window x pass + y pass + solve with the three gradient frames staged ONCE per block.
// strip_window_solve<.., PROD> lets every (channel, half) warp stage the two gradients of its own product: ten staged
// copies of three frames, 151 KB of shared memory for ten warps, which is all an SM then holds (47 % FP64 pipe).  Here the
// block's 5 NQ warps fetch one copy of the RB rows x (32 NQ + K - 1) columns of {dt, dx, dy} together (blocked-transposed
// rows as in StripStage, element-wise cp.async with clamped source addresses = mode='nearest'), and warp (channel ch,
// quarter q) forms its product while it gathers.  The staging is double-buffered across the batch barrier that already
// exists: batch b + 1 is issued at the top of batch b (its buffer was last read by the gathers of b - 1, which every warp
// finished before barrier b - 1), each warp waits for its own copies right before barrier b, and the barrier publishes
// them -- no second barrier, one batch time (thousands of cycles) of prefetch distance.  20 warps (5,5,5,5) at 96 registers.
// The arithmetic is strip_window_solve<.., PROD>'s, operation for operation (bit-identical results).
template <typename T, int K, int NQ>
struct Prod2D {
    static constexpr int NCH = 5, NW = NCH * NQ, TX = 32 * NQ, RL = TX + K - 1, NI = (RL + 7) / 8;
    static constexpr int PITCH = NI | 1;
    static constexpr int ROWSTRIDE = (8 * PITCH + 11) / 16 * 16 + 4;   // = 4 (mod 16): see strip_rowstride
    static constexpr int VOLBUF = kStripRB * ROWSTRIDE, BUF = 3 * VOLBUF;
    static constexpr int PARKROW = strip_parkrow<NCH, NQ>();
    static constexpr size_t smem = (size_t)(2 * BUF + NW * kStripRB * kXwRow + 2 * kStripRB * PARKROW) * sizeof(T);
};

template <typename T, int K, int P, int NQ>
__global__ void __launch_bounds__(5 * NQ * 32, 1) strip_window_solve_2d(const StripArgs<T, K> a) {
    using L = Prod2D<T, K, NQ>;
    constexpr int RB = kStripRB, XB = kStripXB, R = K / 2, NCH = 5, NW = L::NW, TX = L::TX;
    constexpr int PITCH = L::PITCH, ROWSTRIDE = L::ROWSTRIDE, PARKROW = L::PARKROW;
    static_assert(P >= K && P % RB == 0, "unroll period must cover the taps and be a multiple of the batch");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* const stage = reinterpret_cast<T*>(smem_raw);                         // [2][3][RB][ROWSTRIDE]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = warp % NCH, q = warp / NCH;
    T* const xw = stage + 2 * L::BUF + warp * RB * kXwRow;                   // [RB][kXwRow], private to the warp
    T* const park = stage + 2 * L::BUF + NW * RB * kXwRow;                   // [2][RB][PARKROW], shared by the block
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const StripGeom& g = a.g;

    int task = blockIdx.x;
    const int nstrips = (g.n_c + TX - 1) / TX;
    const int strip = task % nstrips; task /= nstrips;
    const int chunk = task % g.n_chunks;
    const int o = task / g.n_chunks;
    const int cs0 = strip * TX;
    const int m0 = chunk * g.chunk;
    const int nout = min(m0 + g.chunk, g.n_m) - m0;
    const int nsteps = (nout + 2 * R + P - 1) / P * P;
    const int64_t plane_off = (int64_t)o * g.stride_o;

    // A quarter-warp fetches a piece of 8 consecutive elements (one per lane) of one row of one gradient.  The 4 NW
    // quarter-warps of the block form RPI rows of NI pieces: which piece (column, destination) a thread fetches is fixed for
    // the whole march, only the row task (gradient, row of the batch) advances -- no division in the loop.
    constexpr int RPI = (4 * NW) / L::NI > 0 ? (4 * NW) / L::NI : 1, NRT = 3 * RB;
    static_assert(L::NI <= 4 * NW, "a row of pieces must fit the block's quarter-warps");
    const int qw = warp * 4 + (lane >> 3);
    const int is_rr = qw / L::NI, is_i = qw - is_rr * L::NI, is_e = 8 * is_i + (lane & 7);
    const bool is_on = is_rr < RPI && (L::RL % 8 == 0 || is_e < L::RL);
    const int is_c = max(0, min(cs0 - R + is_e, g.n_c - 1));                 // clamped source column (mode='nearest')
    const int is_d = (lane & 7) * PITCH + is_i;                              // element offset inside a staged row
    auto issue = [&](int b) {
        if (is_on) {
#pragma unroll 1
            for (int rt = is_rr; rt < NRT; rt += RPI) {
                const int v = rt >> 3, r = rt & 7;
                static_assert(RB == 8, "row task decoding");
                const int m = max(0, min(m0 - R + b * RB + r, g.n_m - 1));
                cp_async_elem<T>(stage_s + (uint32_t)(((b & 1) * L::BUF + v * L::VOLBUF + r * ROWSTRIDE + is_d) * sizeof(T)),
                                 a.in[v] + plane_off + (int64_t)m * g.stride_m + is_c);
            }
        }
        cp_async_commit();
    };

    // channel -> gradient pair {xx,xy,yy,tx,ty}; a.in = {dt, dx, dy}   (calc_flow.py:133-141)
    const int ia = (ch == 2 || ch == 4) ? 2 : 1;
    const int ib = ch == 0 ? 1 : (ch <= 2 ? 2 : 0);
    const int g_r = lane >> 2, g_b = lane & 3;
    const int goff_a = ia * L::VOLBUF + g_r * ROWSTRIDE + g_b + 4 * q;      // element 32 q + 8 g_b of row g_r
    const int goff_d = (ib - ia) * L::VOLBUF;
    T* const m_dst = park + ch * TX + 32 * q + lane;
    T acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) acc[i] = T(0);

    auto solve_batch = [&](int bb) {
#pragma unroll 1
        for (int unit = warp; unit < RB * TX / 32; unit += NW)
            solve_unit_voxels<T, NCH, TX, R>(park, PARKROW, bb, unit, lane, cs0, nout, g.n_c, plane_off, g.stride_m, m0,
                                             a.vx, a.vy, a.vz, a.rel, a.rel_f32);
    };

    issue(0);
    cp_async_wait<0>();
    __syncthreads();
    int b = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int bi = 0; bi < P / RB; ++bi, ++b) {
            issue(b + 1);                                                     // clamped addresses: always valid
            {
                // ---------------- products + gather along x: lane (row g_r, block g_b) -> 8 consecutive outputs
                const T* sa = stage + (b & 1) * L::BUF + goff_a;
                T ga[XB];
#pragma unroll
                for (int i = 0; i < XB; ++i) ga[i] = T(0);
#pragma unroll
                for (int mm = 0; mm < XB + K - 1; ++mm) {
                    T v = sa[(mm & 7) * PITCH + (mm >> 3)];
                    v *= sa[goff_d + (mm & 7) * PITCH + (mm >> 3)];
#pragma unroll
                    for (int i = 0; i < XB; ++i) {
                        const int k = mm - i;
                        if (k >= 0 && k < K) ga[i] = fma(a.f.w[k], v, ga[i]);
                    }
                }
                T* d = xw + g_r * kXwRow + g_b * (XB + 1);
#pragma unroll
                for (int i = 0; i < XB; ++i) d[i] = ga[i];
            }
            __syncwarp();                                                     // gathered rows visible to the warp
            // ---------------- march along y: output j = s - 2R lives in park row j mod RB of buffer b & 1
            {
                T* pk = m_dst + (b & 1) * RB * PARKROW;
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const T res = ring_push<T, K, P, 1>(acc, a.f, xw[r * kXwRow + lane + lane / 8], bi * RB + r);
                    constexpr int kBias = (2 * R + RB - 1) / RB * RB;
                    pk[((r + kBias - 2 * R) % RB) * PARKROW] = res;
                }
            }
            cp_async_wait<0>();                                               // this warp's pieces of batch b + 1 have landed
            __syncthreads();                                                  // batch b parked, batch b + 1 staged, by all warps
            solve_batch(b);                                                   // the RB x TX outputs of batch b
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// fp32 window x pass + y pass + solve with PACKED arithmetic (3D, nine channels).  Warp (channel ch, pair p) owns the two
// adjacent 32-column halves [cs0 + 64 p, +32) and [cs0 + 64 p + 32, +32) of channel ch; a lane's two columns -- one per
// half -- travel through the gather and the march as one f32x2 element: FFMA2 with broadcast tap pairs from uniform
// registers, i.e. half the FMA instructions and half the control flow per voxel of the scalar kernel (which spends its
// time issuing instructions, not computing: 36 % of the FFMA roof).  Staging, park and solve are the scalar kernel's.
template <int K, int NPAIR> constexpr size_t strip_x2_warp_bytes() {
    return (size_t)2 * kStripRB * strip_rowstride<K>() * sizeof(float) + (size_t)kStripRB * kXwRow * sizeof(f32x2);
}
template <int K, int NPAIR> constexpr size_t strip_x2_smem() {
    return 9 * NPAIR * strip_x2_warp_bytes<K, NPAIR>() + (size_t)2 * kStripRB * strip_parkrow<9, 2 * NPAIR>() * sizeof(float);
}

template <int K, int P, int NPAIR, bool SHIFT = false>
__global__ void __launch_bounds__(9 * NPAIR * 32, 1) strip_window_solve_x2(const StripArgs<float, K> a, const Taps<f32x2, K> f2) {
    constexpr int RB = kStripRB, R = K / 2, NCH = 9, NW = NCH * NPAIR, TX = 64 * NPAIR, XB = kStripXB;
    constexpr int PARKROW = strip_parkrow<NCH, 2 * NPAIR>();
    static_assert(P >= K && P % RB == 0, "unroll period must cover the taps and be a multiple of the batch");
    using Stage = StripStage<float, K, 1>;
    constexpr int BUF = Stage::BUF, PITCH = Stage::PITCH, ROWSTRIDE = Stage::ROWSTRIDE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = warp % NCH, pair = warp / NCH;
    unsigned char* wbase = smem_raw + (size_t)warp * strip_x2_warp_bytes<K, NPAIR>();
    float* park = reinterpret_cast<float*>(smem_raw + (size_t)NW * strip_x2_warp_bytes<K, NPAIR>());     // [2][RB][PARKROW]
    f32x2* xw2 = reinterpret_cast<f32x2*>(wbase + 2 * BUF * sizeof(float));                              // [RB][kXwRow]
    const StripGeom& g = a.g;

    int task = blockIdx.x;
    const int nstrips = (g.n_c + TX - 1) / TX;
    const int strip = task % nstrips; task /= nstrips;
    const int chunk = task % g.n_chunks;
    const int o = task / g.n_chunks;
    const int cs0 = strip * TX;
    const int m0 = chunk * g.chunk;
    const int nout = min(m0 + g.chunk, g.n_m) - m0;
    const int nsteps = (nout + 2 * R + P - 1) / P * P;

    Stage st0, st1;                                           // (their own gathered-row buffers are not used)
    st0.init(reinterpret_cast<float*>(wbase), cs0 + 64 * pair, m0 - R, g);
    st1.init(reinterpret_cast<float*>(wbase) + BUF, cs0 + 64 * pair + 32, m0 - R, g);
    st0.src[0] = st1.src[0] = a.in[0] + (int64_t)o * g.stride_o + (int64_t)ch * g.vol;
    float* const m_dst = park + ch * TX + 64 * pair + lane;

    // The solve units of a batch are dealt to the four schedulers (warp % 4) so that FFMA2 work + fp64 solve work per
    // scheduler is even: in this kernel the solve is HALF of the instructions (ncu: 353 per voxel incl. conversions,
    // addressing and stores, against 365 for gather + march), so a scheduler with five marching warps takes 7 of the 32
    // units and one with four takes 9 (round robin over the warps gave 9,9,7,7: the wrong way round).
    constexpr int NU = RB * TX / 32;
    constexpr int C0 = NPAIR == 2 ? 7 : 3, C1 = NPAIR == 2 ? 7 : 4, C2 = NPAIR == 2 ? 9 : 5;
    static_assert(NPAIR == 1 || NPAIR == 2, "unit table");
    auto solve_batch = [&](int bb) {
        const int w = threadIdx.x >> 5, sch = w & 3, nws = (NW - sch + 3) >> 2;      // (derived here: not live across the marches)
        const int u_begin = (sch == 0 ? 0 : (sch == 1 ? C0 : (sch == 2 ? C0 + C1 : C0 + C1 + C2))) + (w >> 2);
        const int u_end = sch == 0 ? C0 : (sch == 1 ? C0 + C1 : (sch == 2 ? C0 + C1 + C2 : NU));
#pragma unroll 1
        for (int unit = u_begin; unit < u_end; unit += nws)
            solve_unit_voxels<float, NCH, TX, R>(park, PARKROW, bb, unit, lane, cs0, nout, g.n_c, (int64_t)o * g.stride_o, g.stride_m, m0,
                                                 a.vx, a.vy, a.vz, a.rel, 0);
    };

    st0.issue(0);
    st1.issue(0);
    // gather along x, both halves at once: lane (row g_r, block g_b) -> 8 consecutive outputs
    auto gather_x = [&]() {
        const int g_r = lane >> 2, g_b = lane & 3;
        const float* r0 = st0.rows + g_r * ROWSTRIDE + g_b;
        const float* r1 = st1.rows + g_r * ROWSTRIDE + g_b;
        f32x2 ga[XB];
#pragma unroll
        for (int mm = 0; mm < XB + K - 1; ++mm) {
            const f32x2 v(r0[(mm & 7) * PITCH + (mm >> 3)], r1[(mm & 7) * PITCH + (mm >> 3)]);
#pragma unroll
            for (int i = 0; i < XB; ++i) {
                const int k = mm - i;
                if (k >= 0 && k < K) {
                    const f32x2 w = f2.w[k > R ? K - 1 - k : k];      // symmetric window: half of the taps
                    if (k == 0) mul_acc(ga[i], w, v); else fma_acc(ga[i], w, v);
                }
            }
        }
        f32x2* d = xw2 + g_r * kXwRow + g_b * (XB + 1);
#pragma unroll
        for (int i = 0; i < XB; ++i) d[i] = ga[i];
    };
    if constexpr (SHIFT) {
        // y march on the shifting ring (kernels_march.cuh shift_push): K packed accumulators, one copy of the batch
        f32x2 accs[K];
#pragma unroll
        for (int i = 0; i < K; ++i) accs[i] = f32x2(0);
        const int nb = (nout + 2 * R + RB - 1) / RB;
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            cp_async_wait<0>();
            __syncwarp();
            gather_x();
            __syncwarp();
            st0.issue(b + 1);
            st1.issue(b + 1);
            float* pk = m_dst + (b & 1) * RB * PARKROW;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const f32x2 res = shift_push<f32x2, K, 1>(accs, f2, xw2[r * kXwRow + lane + lane / 8]);
                constexpr int kBias = (2 * R + RB - 1) / RB * RB;
                float* q = pk + ((r + kBias - 2 * R) % RB) * PARKROW;
                q[0] = res.lo();
                q[32] = res.hi();
            }
            __syncthreads();
            solve_batch(b);
        }
        cp_async_wait<0>();
        return;
    }
    f32x2 acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) acc[i] = f32x2(0);
    int b = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int bi = 0; bi < P / RB; ++bi, ++b) {
            cp_async_wait<0>();
            __syncwarp();                                                     // this warp's rows of batch b have landed
            gather_x();
            __syncwarp();                                                     // gathered rows visible; input rows free
            st0.issue(b + 1);                                                 // prefetch (clamped addresses: always valid)
            st1.issue(b + 1);
            // ---------------- march along y: output j = s - 2R lives in park row j mod RB of buffer b & 1
            {
                float* pk = m_dst + (b & 1) * RB * PARKROW;
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const f32x2 res = ring_push<f32x2, K, P, 1>(acc, f2, xw2[r * kXwRow + lane + lane / 8], bi * RB + r);
                    constexpr int kBias = (2 * R + RB - 1) / RB * RB;
                    float* q = pk + ((r + kBias - 2 * R) % RB) * PARKROW;
                    q[0] = res.lo();
                    q[32] = res.hi();
                }
            }
            __syncthreads();                                                  // one batch of outputs parked by all channels
            solve_batch(b);
        }
    }
    cp_async_wait<0>();
}

}  // namespace of3d
