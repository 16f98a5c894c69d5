// Shared declarations for libof3d (sm_100a). Internal header.
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/of3d.h"

namespace of3d {

constexpr int kMaxTaps = 257;   // radius <= 128  (sigma <= 42)
constexpr int kMaxFrames = 129; // temporal taps (tSig <= 21)

// One 1-D filter passed BY VALUE as a kernel parameter: it lands in constant bank 0, so
// statically indexed taps become immediate constant operands of DFMA/FFMA and there is no
// shared mutable __constant__ state between contexts.
template <typename T>
struct Filt {
    int n;      // number of taps (odd)
    int sym;    // +1 symmetric, -1 antisymmetric, 0 neither (scipy's NI_Correlate1D test)
    T w[kMaxTaps];
};

struct FramePtrs {
    const void* p[kMaxFrames];
};

void set_error(const std::string& msg);
const char* cuda_err_name(cudaError_t e);

#define OF3D_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::of3d::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));     \
            return (_e == cudaErrorMemoryAllocation) ? OF3D_ERR_NOMEM : OF3D_ERR_CUDA;        \
        }                                                                                     \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a = 512) { return (x + a - 1) / a * a; }

struct Shape {
    int ndim;            // 2 or 3
    int64_t nz, ny, nx;  // nz == 1 for 2D
    int64_t n() const { return nz * ny * nx; }
};

}  // namespace of3d

// The opaque context of the C ABI.
struct of3d_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // the stream every kernel of the context runs on
    cudaStream_t own_stream = nullptr;   // created with the context; `stream` is this one unless the caller lent its own
    char* ws = nullptr;   // device workspace
    size_t ws_cap = 0;
    size_t ws_off = 0;    // bump pointer (reset every call)
    int async = 0;
    int64_t launches = 0;
    int sm_count = 148;
    // device-resident input window fed frame by frame (of3d_window_upload / of3d_window_flow)
    cudaStream_t s_up = nullptr, s_dn = nullptr;
    cudaEvent_t ev_up = nullptr;
    char* win = nullptr;                                // [win_n frames][slot]
    size_t win_cap = 0, win_frame = 0;
    int win_n = 0;
    struct Part { size_t off; cudaEvent_t ev; };       // uploaded pieces in shipping order (byte offset in a frame)
    std::vector<Part> parts;
    std::vector<cudaEvent_t> part_pool;
    char* pipe = nullptr;                               // ic, dt0 and two sets of slab outputs of the pipelined host call
    size_t pipe_cap = 0;
    cudaEvent_t ev_c = nullptr, ev_dn[2] = {nullptr, nullptr};
    // z-slab sharding: NCCL communicator (loaded at run time) and the halo-exchange stream
    void* comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    cudaStream_t s_comm = nullptr;
    cudaEvent_t ev_halo = nullptr, ev_ready = nullptr;
    bool halo_pending = false;                          // an exchange is in flight: slab calls wait for it where they need the halo
    // optional per-stage device timing (of3d_set_profile): every launch is bracketed by events on `stream`
    int profile = 0;
    std::vector<cudaEvent_t> ev_pool;                 // recycled events
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector<Span> spans;                          // recorded, not yet resolved
    double stage_ms[OF3D_N_STAGES] = {};
    int64_t stage_launches[OF3D_N_STAGES] = {};
};

namespace of3d {

// Brackets the launches issued during its lifetime with two events when the context is profiling.
struct StageScope {
    of3d_ctx* c;
    int stage;
    cudaEvent_t a = nullptr;
    static cudaEvent_t take(of3d_ctx* c) {
        cudaEvent_t e = nullptr;
        if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    StageScope(of3d_ctx* ctx, int st) : c(ctx), stage(st) {
        if (c->profile) { a = take(c); cudaEventRecord(a, c->stream); }
    }
    ~StageScope() {
        if (a) { cudaEvent_t b = take(c); cudaEventRecord(b, c->stream); c->spans.push_back({stage, a, b}); }
    }
};

template <typename P>
static inline P* ws_take(of3d_ctx* c, size_t count) {
    P* p = reinterpret_cast<P*>(c->ws + c->ws_off);
    c->ws_off += align_up(count * sizeof(P));
    return p;
}

// Marching-kernel pipeline (pipeline_fast.cuh): returns OF3D_OK, an error, or kNotSupported when the
// tap counts have no specialised instantiation (the caller then runs the generic pipeline).
constexpr int kNotSupported = 1;
// Flow of the planes [own_lo, own_lo + own_n) of the volume s (3D; own_lo = 0, own_n = 1 in 2D); the outputs hold own_n
// planes.  Temporal stage: raw frames `fp` of `in_dtype` (3D only, when fused_temporal_ok) or (ic, dt0) volumes.
template <typename T>
int run_fast(of3d_ctx* c, const Shape& s, const FramePtrs* fp, int in_dtype, const T* ic, const T* dt0, const of3d_taps* t,
             int64_t own_lo, int64_t own_n, T* vx, T* vy, T* vz, T* rel, int rel_f32);

// Tap-count classes with a specialised marching instantiation (shorter filters are zero-padded to the class)
static inline int spatial_class(const of3d_taps* t) {
    const int kr = t->nD > t->nG ? t->nD : t->nG, ks = t->nS;
    if (kr <= 7 && ks <= 3) return 0;
    if (kr <= 13 && ks <= 5) return 1;
    if (kr <= 19 && ks <= 7) return 2;
    if (kr <= 25 && ks <= 7) return 3;
    return -1;
}
static inline int window_class(const of3d_taps* t) {
    const int ks[6] = {7, 13, 19, 25, 37, 49};
    for (int i = 0; i < 6; ++i)
        if (t->nW <= ks[i]) return i;
    return -1;
}
// The marching kernels read only half of every (anti)symmetric filter: S, G, W symmetric and D antisymmetric, bit for bit
// (true of the reference's sampled Gaussians, calc_flow.py:238-266; arbitrary taps run on the generic kernels)
static inline bool taps_symmetric(const double* w, int n, double sign) {
    for (int i = 0; i < n / 2; ++i)
        if (w[n - 1 - i] != sign * w[i]) return false;
    return true;
}
static inline bool fast_supported(const of3d_taps* t) {
    return spatial_class(t) >= 0 && window_class(t) >= 0 && taps_symmetric(t->S, t->nS, 1.0) && taps_symmetric(t->G, t->nG, 1.0) &&
           taps_symmetric(t->W, t->nW, 1.0) && taps_symmetric(t->D, t->nD, -1.0);
}
// workspace volumes of compute type used by run_fast on a whole volume (excluding ic and dt0)
static inline int fast_volumes(int ndim) { return ndim == 3 ? 9 + 4 : 3; }
// The z march can read the raw frames itself (kernels_tz.cuh) when they are 8/16-bit unsigned integers, every frame and row is
// 16-byte aligned, the window has at most kFusedMaxFrames frames and T is antisymmetric bit for bit with a zero centre
constexpr int kFusedMaxFrames = 24;
static inline bool fused_temporal_ok(const of3d_taps* t, const FramePtrs& fp, int in_dtype, int64_t nx) {
    const int sz = in_dtype == OF3D_U8 ? 1 : (in_dtype == OF3D_U16 ? 2 : 0);
    if (!sz || t->nT > kFusedMaxFrames || (nx * sz) % 16 != 0 || !taps_symmetric(t->T, t->nT, -1.0) || t->T[t->nT / 2] != 0.0) return false;
    for (int k = 0; k < t->nT; ++k)
        if (reinterpret_cast<uintptr_t>(fp.p[k]) % 16) return false;
    return true;
}

}  // namespace of3d
