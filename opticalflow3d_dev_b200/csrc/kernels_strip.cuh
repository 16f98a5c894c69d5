// Plane-strip kernel: the last TWO window passes and the per-voxel solve in one launch.  Internal header.
//
// A block owns a strip of TX = 64 positions of the contiguous axis ("c") of one plane and marches along the
// other in-plane axis ("m") in batches of RB = 8 rows.  It has NCH * 2 warps: warp (channel ch, half h) owns the
// 32 columns [cs0 + 32 h, cs0 + 32 h + 32) of channel ch, and everything up to the solve is private to the warp
// (only __syncwarp):
//   load    cp.async of the warp's RB rows x (32 + K - 1) columns into its shared-memory rows (clamp-to-edge on
//           the source address), issued one batch ahead;
//   gather  window along c: lane (row r = l/4, block b = l%4) produces 8 consecutive outputs of row r from the row
//           in shared memory -- 8 accumulators, 8 + K - 1 LDS, 8*K FMAs, all addresses and taps static.  Rows are
//           stored "blocked-transposed" (element e at (e % 8) * PITCH + e / 8) so that lanes read consecutive words;
//   march   window along m: lane = column; the gathered value of each of the RB rows is scattered into the K
//           register accumulators (same statically rotated ring as kernels_march.cuh); completed sums are parked in
//           shared memory for the whole block.
// Then ONE block barrier per batch, and every thread solves one voxel of the RB x TX batch (calc_flow.py:337-357)
// and stores vx, vy, vz, rel.  The park is double-buffered, so warps may run up to a batch apart: the latency-bound
// solve of one warp overlaps the FMA-bound gather/march of others.  The 9 window sums exist only in registers and
// shared memory: compared with separate x and y passes this removes a 72 B/voxel write and a 72 B/voxel read of HBM
// (fp64) and one launch.
//
// With TR the input is in T layout (z, x, y): c = y, m = x, and the outputs are written transposed, i.e. in
// N layout (z, y, x) as the C ABI requires.
#pragma once
#include "common.cuh"
#include "kernels_march.cuh"
#include "solve.cuh"

namespace of3d {

constexpr int kStripTX = 64;   // strip width along the contiguous axis (two warps of 32 columns per channel)
constexpr int kStripRB = 8;    // rows per batch
constexpr int kStripXB = 8;    // outputs per lane in the gather phase

template <int K> constexpr int strip_rowlen() { return 32 + K - 1; }
template <int K> constexpr int strip_nblk() { return (strip_rowlen<K>() + 7) / 8; }
template <int K> constexpr int strip_pitch() { return strip_nblk<K>() | 1; }
// row stride = 4 (mod 16) elements: the four rows a half-warp reads in the gather fall on disjoint banks
template <int K> constexpr int strip_rowstride() { return (8 * strip_pitch<K>() + 11) / 16 * 16 + 4; }
constexpr int kXwRow = 32 + 4;                                               // one pad element every 8 columns
template <int NCH> constexpr int strip_parkrow() { return NCH * kStripTX + 2; }
template <int K> constexpr int strip_warp_elems() { return kStripRB * strip_rowstride<K>() + kStripRB * kXwRow; }

template <typename T, int K, int NCH>
constexpr size_t strip_smem() {
    return (size_t)(2 * NCH * strip_warp_elems<K>() + 2 * kStripRB * strip_parkrow<NCH>()) * sizeof(T);
}

template <typename T, int K>
struct StripArgs {
    Taps<T, K> f;
    const T* in;            // channel-major volumes
    T* vx; T* vy; T* vz; T* rel;
    int64_t vol;            // elements per channel
    int64_t stride_m;       // element stride of the march axis (input)
    int64_t stride_o;       // element stride of the remaining axis (input and output)
    int n_c, n_m, n_o;      // extents: contiguous, march, other
    int chunk, n_chunks;    // outputs per block along m
};

template <typename T, int K, int P, int NCH, bool TR>
__global__ void __launch_bounds__(NCH * kStripTX, 1) strip_window_solve(const StripArgs<T, K> a) {
    constexpr int TX = kStripTX, RB = kStripRB, XB = kStripXB, R = K / 2;
    constexpr int ROWLEN = strip_rowlen<K>(), PITCH = strip_pitch<K>(), ROWSTRIDE = strip_rowstride<K>();
    constexpr int PARKROW = strip_parkrow<NCH>();
    constexpr int NLOAD = (ROWLEN + 7) / 8;                                  // cp.async per lane per row pass
    static_assert(P >= K && P % RB == 0, "unroll period must cover the taps and be a multiple of the batch");
    static_assert(ROWSTRIDE >= 8 * PITCH, "row stride too small");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ch = warp % NCH, half = warp / NCH;
    T* rows = smem + warp * strip_warp_elems<K>();                           // [RB][ROWSTRIDE], private to the warp
    T* xw = rows + RB * ROWSTRIDE;                                           // [RB][kXwRow], private to the warp
    T* park = smem + 2 * NCH * strip_warp_elems<K>();                        // [2][RB][PARKROW], shared by the block
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);

    int task = blockIdx.x;
    const int nstrips = (a.n_c + TX - 1) / TX;
    const int strip = task % nstrips; task /= nstrips;
    const int chunk = task % a.n_chunks;
    const int o = task / a.n_chunks;
    const int cs0 = strip * TX;                                              // first column of the block's strip
    const int cw0 = cs0 + 32 * half;                                         // first column of this warp
    const int m0 = chunk * a.chunk;
    const int m1 = min(m0 + a.chunk, a.n_m);
    const int nout = m1 - m0;
    const int nsteps = (nout + 2 * R + P - 1) / P * P;                       // whole unroll periods
    const T* in_c = a.in + (int64_t)o * a.stride_o + (int64_t)ch * a.vol;

    // ---- loader: lane (l8 = lane % 8, rr = lane / 8) fetches elements e = l8 + 8 i of rows rr and rr + 4; element e
    // lives at (e % 8) * PITCH + e / 8 = l8 * PITCH + i.  Clamp-to-edge along c on the source column, along m on the row.
    const int l8 = lane & 7, lrr = lane >> 3;
    const int l_c0 = cw0 - R + l8;
    const int ncm1 = a.n_c - 1, nmm1 = a.n_m - 1;
    auto issue_batch = [&](int b) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int r = lrr + 4 * p;
            int m = m0 - R + b * RB + r;
            m = max(0, min(m, nmm1));
            const T* row = in_c + (int64_t)m * a.stride_m;
            const uint32_t dst = rows_s + (uint32_t)((r * ROWSTRIDE + l8 * PITCH) * sizeof(T));
#pragma unroll
            for (int i = 0; i < NLOAD; ++i) {
                if (ROWLEN % 8 == 0 || l8 + 8 * i < ROWLEN) {
                    const int c = max(0, min(l_c0 + 8 * i, ncm1));
                    cp_async_elem<T>(dst + (uint32_t)(i * sizeof(T)), row + c);
                }
            }
        }
        cp_async_commit();
    };

    // gather role: row g_r, block g_b of 8 outputs;  march role: column `lane`;  solve role: one voxel of RB x TX
    const int g_r = lane >> 2, g_b = lane & 3;
    const T* g_src = rows + g_r * ROWSTRIDE + g_b;
    T* g_dst = xw + g_r * kXwRow + g_b * (XB + 1);
    const T* m_src = xw + lane + lane / 8;
    T* m_dst = park + ch * TX + 32 * half + lane;
    // Solve work is handed out in units of 32 voxels (RB * TX / 32 = 16 units per batch).  A block has 18 warps on 4
    // schedulers (warp % 4): two schedulers run 5 window warps, two run 4.  The lighter schedulers take more solve
    // units (5 each) than the heavier ones (3 each), which evens out the FP64 work per scheduler.
    constexpr int NUNITS = RB * TX / 32;
    int su[2] = {-1, -1};
    if (NCH == 9) {
        const int q = warp >> 2, sm = warp & 3;                               // q-th warp of scheduler sm
        if (sm >= 2) { su[0] = (sm - 2) * 5 + q; if (q == 0) su[1] = (sm - 2) * 5 + 4; }   // 4 warps, 5 units
        else if (q < 3) su[0] = 10 + sm * 3 + q;                              // 5 warps, 3 units
    } else {
        su[0] = warp < NUNITS ? warp : -1;
        su[1] = warp + 2 * NCH < NUNITS ? warp + 2 * NCH : -1;
    }
    static_assert(NCH != 9 || NUNITS == 16, "unit table assumes 16 units");
    T acc[P];
#pragma unroll
    for (int i = 0; i < P; ++i) acc[i] = T(0);

    issue_batch(0);
    int b = 0;
#pragma unroll 1
    for (int s0 = 0; s0 < nsteps; s0 += P) {
#pragma unroll
        for (int bi = 0; bi < P / RB; ++bi, ++b) {
            cp_async_wait<0>();
            __syncwarp();                                                     // this warp's rows of batch b have landed
            // ---------------- gather along c
            {
                T ga[XB];
#pragma unroll
                for (int i = 0; i < XB; ++i) ga[i] = T(0);
#pragma unroll
                for (int mm = 0; mm < XB + K - 1; ++mm) {
                    const T v = g_src[(mm & 7) * PITCH + (mm >> 3)];
#pragma unroll
                    for (int i = 0; i < XB; ++i) {
                        const int k = mm - i;
                        if (k >= 0 && k < K) ga[i] = fma(a.f.w[k], v, ga[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < XB; ++i) g_dst[i] = ga[i];
            }
            __syncwarp();                                                     // gathered rows visible; input rows free
            issue_batch(b + 1);                                               // prefetch (clamped addresses: always valid)
            // ---------------- march along m: output j = s - 2R lives in park row j mod RB of buffer b & 1
            {
                T* pk = m_dst + (b & 1) * RB * PARKROW;
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const T v = m_src[r * kXwRow];
                    const T res = ring_push<T, K, P>(acc, a.f, v, bi * RB + r);
                    constexpr int kBias = (2 * R + RB - 1) / RB * RB;
                    pk[((r + kBias - 2 * R) % RB) * PARKROW] = res;
                }
            }
            __syncthreads();                                                  // one batch of outputs parked by all channels
            // ---------------- solve: outputs j0 .. j0 + RB - 1 of batch b
#pragma unroll
            for (int ui = 0; ui < 2; ++ui) {
                if (su[ui] < 0) continue;                                     // warp-uniform
                const int v = su[ui] * 32 + lane;                             // voxel of the RB x TX batch
                const int s_i = TR ? (v % RB) : (v / TX);                     // output row within the batch
                const int s_col = TR ? (v / RB) : (v % TX);
                const int s_c = cs0 + s_col;
                const int j = b * RB - 2 * R + s_i;
                if (j >= 0 && j < nout && s_c < a.n_c) {
                    const int prow = ((j % RB) + RB) % RB;
                    const T* qv = park + ((b & 1) * RB + prow) * PARKROW + s_col;
                    const int64_t idx = TR ? ((int64_t)o * a.stride_o + (int64_t)s_c * a.n_m + (m0 + j))
                                           : ((int64_t)o * a.stride_o + (int64_t)(m0 + j) * a.stride_m + s_c);
                    if (NCH == 9) {
#ifdef OF3D_EXP_NOSOLVE
                        Flow3 rr; rr.vx = qv[0] + qv[TX] + qv[2 * TX]; rr.vy = qv[3 * TX] + qv[4 * TX]; rr.vz = qv[5 * TX] + qv[6 * TX]; rr.rel = qv[7 * TX] + qv[8 * TX];
#else
                        const Flow3 rr = solve3<false>((double)qv[0], (double)qv[TX], (double)qv[2 * TX], (double)qv[3 * TX],
                                                       (double)qv[4 * TX], (double)qv[5 * TX], (double)qv[6 * TX],
                                                       (double)qv[7 * TX], (double)qv[8 * TX]);
#endif
                        a.vx[idx] = (T)rr.vx; a.vy[idx] = (T)rr.vy; a.vz[idx] = (T)rr.vz; a.rel[idx] = (T)rr.rel;
                    } else {
                        const Flow2 rr = solve2<false>((double)qv[0], (double)qv[TX], (double)qv[2 * TX], (double)qv[3 * TX],
                                                       (double)qv[4 * TX]);
                        a.vx[idx] = (T)rr.vx; a.vy[idx] = (T)rr.vy; a.rel[idx] = (T)rr.rel;
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
}

}  // namespace of3d
