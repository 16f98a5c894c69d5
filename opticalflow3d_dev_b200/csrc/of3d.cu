// libof3d: C ABI + host-side orchestration of the Lucas-Kanade pipeline on one B200.
// See include/of3d.h for the contract and DESIGN.md for the kernel plan.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <cstdlib>
#include <vector>
#include <dlfcn.h>

#include "common.cuh"
#include "kernels_analysis.cuh"
#include "kernels_generic.cuh"
#include "kernels_synth.cuh"
#include "solve.cuh"

namespace of3d {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

}  // namespace of3d

using namespace of3d;

namespace of3d {

static size_t dtype_size(int dt) {
    switch (dt) {
        case OF3D_U8: return 1;
        case OF3D_U16: case OF3D_I16: return 2;
        case OF3D_F32: case OF3D_I32: case OF3D_U32: return 4;
        case OF3D_F64: return 8;
        default: return 0;
    }
}

// Scipy's symmetry test (ni_filters.c): odd length and |w[r+i] -/+ w[r-i]| <= DBL_EPSILON
static int tap_symmetry(const double* w, int n) {
    if (!(n & 1)) return 0;
    const int r = n / 2;
    bool sym = true, asym = true;
    for (int i = 1; i <= r; ++i) {
        if (std::fabs(w[r + i] - w[r - i]) > 2.220446049250313e-16) sym = false;
        if (std::fabs(w[r + i] + w[r - i]) > 2.220446049250313e-16) asym = false;
    }
    return sym ? 1 : (asym ? -1 : 0);
}

template <typename T>
static Filt<T> make_filt(const double* w, int n) {
    Filt<T> f;
    f.n = n;
    f.sym = tap_symmetry(w, n);
    for (int i = 0; i < kMaxTaps; ++i) f.w[i] = i < n ? (T)w[i] : T(0);  // rounded once to the compute type
    return f;
}

static int grid_for(const of3d_ctx* c, int64_t n, int block = 256) {
    int64_t g = ceil_div(n, block);
    int64_t cap = (int64_t)c->sm_count * 32;  // grid-stride beyond this
    return (int)std::max<int64_t>(1, std::min(g, cap));
}

// ---------------------------------------------------------------------------------------------
// workspace
static int ws_ensure(of3d_ctx* c, size_t bytes) {
    if (bytes <= c->ws_cap) return OF3D_OK;
    OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->ws) { OF3D_CUDA_TRY(cudaFree(c->ws)); c->ws = nullptr; c->ws_cap = 0; }
    cudaError_t e = cudaMalloc(&c->ws, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("device workspace allocation of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
        return OF3D_ERR_NOMEM;
    }
    c->ws_cap = bytes;
    return OF3D_OK;
}

// Volumes of compute type needed by the generic pipeline (see GenericPipe::run_spatial), excluding ic and dt0
static int generic_volumes(int ndim) { return ndim == 3 ? 2 + 4 + 1 + 9 : 2 + 3 + 1 + 5; }

// Device workspace of one call.  nz planes are given, own_n of them are produced (3D ranges: z-slab sharding, slab
// pipeline).  fast / fused say which pipeline will run (upper bound over both when unknown: fast = -1).
static size_t plan_bytes(int ndim, int64_t kt, int64_t nz, int64_t own_n, int64_t plane, int nw, int in_dtype, int precision, int in_mem,
                         int out_mem, int fast, bool fused, bool rel_f32) {
    const size_t ts = precision == OF3D_FP32 ? 4 : 8;
    const size_t n = (size_t)nz * plane, vol = align_up(n * ts);
    size_t fast_b = 0, gen_b = 0;
    if (ndim == 3) {
        const size_t ng = align_up((size_t)std::min<int64_t>(nz, own_n + 2 * (nw / 2)) * plane * ts);
        fast_b = std::max(3 * ng, align_up(9 * (size_t)own_n * plane * ts)) + 4 * ng + (fused ? 0 : 2 * vol);
    } else {
        fast_b = 5 * vol;
    }
    gen_b = (size_t)(2 + generic_volumes(ndim)) * vol + (own_n != nz ? 4 * vol : 0) + (rel_f32 ? align_up(n * 8) : 0);
    size_t b = fast < 0 ? std::max(fast_b, gen_b) : (fast ? fast_b : gen_b);
    if (in_mem == OF3D_HOST) b += (size_t)kt * align_up(n * dtype_size(in_dtype));
    if (out_mem == OF3D_HOST) b += (size_t)(ndim + 1) * align_up((size_t)own_n * plane * ts);
    return b + 8192;
}

// ---------------------------------------------------------------------------------------------
// generic pipeline: reference pass order (y -> x -> z), one pass per launch
template <typename T, bool EXACT>
struct GenericPipe {
    of3d_ctx* c;
    Shape s;
    Filt<T> fD, fS, fG, fT, fW;
    bool tpaired;       // T antisymmetric bit for bit with a zero centre: integer frames use the paired temporal form

    void corr(const T* in, T* out, int axis /*0=z,1=y,2=x*/, const Filt<T>& f) {
        const int64_t n = s.n();
        const int64_t len = axis == 0 ? s.nz : (axis == 1 ? s.ny : s.nx);
        const int64_t stride = axis == 0 ? s.ny * s.nx : (axis == 1 ? s.nx : 1);
        StageScope span(c, OF3D_STAGE_GENERIC);
        corr_axis_generic<T, EXACT><<<grid_for(c, n), 256, 0, c->stream>>>(in, out, n, len, stride, f);
        c->launches++;
    }
    // out = F_z F_x F_y in   (z skipped in 2D); tmp1/tmp2 scratch
    void chain(const T* in, T* out, T* t1, T* t2, const Filt<T>& fy, const Filt<T>& fx, const Filt<T>& fz) {
        corr(in, t1, 1, fy);
        if (s.ndim == 3) { corr(t1, t2, 2, fx); corr(t2, out, 0, fz); }
        else corr(t1, out, 2, fx);
    }

    template <typename Tin>
    void temporal(const FramePtrs& fp, T* ic, T* dt0) {
        StageScope span(c, OF3D_STAGE_TEMPORAL);
        const int64_t n = s.n();
        const int paired = !EXACT && tpaired;
        if (!EXACT) {
            // 16-byte vector path when every frame (and the outputs) is 16-byte aligned; scalar kernel for the tail
            constexpr int VEC = 16 / sizeof(Tin);
            bool aligned = (reinterpret_cast<uintptr_t>(ic) | reinterpret_cast<uintptr_t>(dt0)) % 16 == 0;
            for (int k = 0; k < fT.n; ++k) aligned = aligned && reinterpret_cast<uintptr_t>(fp.p[k]) % 16 == 0;
            const int64_t nvec = aligned ? n / VEC : 0;
            if (nvec > 0) {
                temporal_vec<Tin, T><<<grid_for(c, nvec), 256, 0, c->stream>>>(fp, fT, ic, dt0, nvec, paired);
                c->launches++;
                const int64_t done = nvec * VEC;
                if (done < n) {
                    FramePtrs tail = fp;
                    for (int k = 0; k < fT.n; ++k) tail.p[k] = reinterpret_cast<const Tin*>(fp.p[k]) + done;
                    temporal_generic<Tin, T, false><<<grid_for(c, n - done), 256, 0, c->stream>>>(tail, fT, ic + done, dt0 + done, n - done, paired);
                    c->launches++;
                }
                return;
            }
        }
        temporal_generic<Tin, T, EXACT><<<grid_for(c, n), 256, 0, c->stream>>>(fp, fT, ic, dt0, n, paired);
        c->launches++;
    }

    int run_temporal(const FramePtrs& fp, int in_dtype, T* ic, T* dt0) {
        switch (in_dtype) {
            case OF3D_U8: temporal<uint8_t>(fp, ic, dt0); break;
            case OF3D_U16: temporal<uint16_t>(fp, ic, dt0); break;
            case OF3D_I16: temporal<int16_t>(fp, ic, dt0); break;
            case OF3D_F32: temporal<float>(fp, ic, dt0); break;
            case OF3D_F64: temporal<double>(fp, ic, dt0); break;
            case OF3D_I32: temporal<int32_t>(fp, ic, dt0); break;
            case OF3D_U32: temporal<uint32_t>(fp, ic, dt0); break;
            default: set_error("unsupported input dtype"); return OF3D_ERR_ARG;
        }
        return OF3D_OK;
    }

    // everything after the temporal stage: ic = widened centre frame, dt0 = temporal derivative
    int run_spatial(const T* ic, const T* dt0, T* vx, T* vy, T* vz, T* rel) {
        const int64_t n = s.n();
        T* t1 = ws_take<T>(c, n);
        T* t2 = ws_take<T>(c, n);
        T* dt = ws_take<T>(c, n);
        T* dx = ws_take<T>(c, n);
        T* dy = ws_take<T>(c, n);
        T* dz = s.ndim == 3 ? ws_take<T>(c, n) : nullptr;
        T* prod = ws_take<T>(c, n);
        const int nch = s.ndim == 3 ? 9 : 5;
        T* w = ws_take<T>(c, (size_t)nch * n);
        chain(dt0, dt, t1, t2, fG, fG, fG);          // calc_flow.py:279 / 116
        chain(ic, dy, t1, t2, fD, fS, fS);           // :282 / 119
        chain(ic, dx, t1, t2, fS, fD, fS);           // :285 / 122
        if (s.ndim == 3) chain(ic, dz, t1, t2, fS, fS, fD);  // :288
        // products + window, channel order {xx,xy,xz,yy,yz,zz,tx,ty,tz} / {xx,xy,yy,tx,ty}
        const T* pa[9]; const T* pb[9];
        if (s.ndim == 3) {
            const T* A[9] = {dx, dx, dx, dy, dy, dz, dx, dy, dz};
            const T* B[9] = {dx, dy, dz, dy, dz, dz, dt, dt, dt};
            for (int i = 0; i < 9; ++i) { pa[i] = A[i]; pb[i] = B[i]; }
        } else {
            const T* A[5] = {dx, dx, dy, dx, dy};
            const T* B[5] = {dx, dy, dy, dt, dt};
            for (int i = 0; i < 5; ++i) { pa[i] = A[i]; pb[i] = B[i]; }
        }
        for (int ch = 0; ch < nch; ++ch) {
            StageScope span(c, OF3D_STAGE_GENERIC);
            product_generic<T><<<grid_for(c, n), 256, 0, c->stream>>>(pa[ch], pb[ch], prod, n);
            c->launches++;
            chain(prod, w + (size_t)ch * n, t1, t2, fW, fW, fW);  // :300-313 / 133-141
        }
        StageScope span(c, OF3D_STAGE_GENERIC);
        if (s.ndim == 3) solve3_generic<T, EXACT><<<grid_for(c, n), 256, 0, c->stream>>>(w, n, vx, vy, vz, rel);
        else solve2_generic<T, EXACT><<<grid_for(c, n), 256, 0, c->stream>>>(w, n, vx, vy, rel);
        c->launches++;
        return OF3D_OK;
    }
};

static int check_taps(const of3d_taps* t) {
    if (!t) { set_error("taps is null"); return OF3D_ERR_ARG; }
    const double* p[5] = {t->D, t->S, t->G, t->T, t->W};
    const int n[5] = {t->nD, t->nS, t->nG, t->nT, t->nW};
    const char* nm = "DSGTW";
    for (int i = 0; i < 5; ++i) {
        if (!p[i] || n[i] < 1 || !(n[i] & 1)) { set_error(std::string("tap vector ") + nm[i] + " must be non-null with odd length"); return OF3D_ERR_ARG; }
        const int lim = i == 3 ? kMaxFrames : kMaxTaps;
        if (n[i] > lim) { set_error(std::string("tap vector ") + nm[i] + " is longer than the supported " + std::to_string(lim)); return OF3D_ERR_ARG; }
    }
    return OF3D_OK;
}

// fp != nullptr: the temporal stage runs on the frames (into (ic_out, dt0_out) when given);
// fp == nullptr: (ic_in, dt0_in) are given.  spatial == false stops after the temporal stage.
// 3D: the flow of the planes [own_lo, own_lo + own_n) is written to outputs of own_n planes.
template <typename T, bool EXACT>
static int run_pipe(of3d_ctx* c, const Shape& s, const FramePtrs* fp, int in_dtype, const of3d_taps* t, unsigned flags,
                    const T* ic_in, const T* dt0_in, T* ic_out, T* dt0_out, bool spatial, int64_t own_lo, int64_t own_n,
                    T* vx, T* vy, T* vz, T* rel, bool rel_f32) {
    GenericPipe<T, EXACT> g{c, s, make_filt<T>(t->D, t->nD), make_filt<T>(t->S, t->nS), make_filt<T>(t->G, t->nG),
                            make_filt<T>(t->T, t->nT), make_filt<T>(t->W, t->nW),
                            taps_symmetric(t->T, t->nT, -1.0) && t->T[t->nT / 2] == 0.0};
    const bool fast = !EXACT && !(flags & OF3D_FLAG_GENERIC) && fast_supported(t);
    // marching kernels with the temporal derivative fused into the z march (kernels_tz.cuh)
    if (fast && spatial && fp && !ic_out && s.ndim == 3 && fused_temporal_ok(t, *fp, in_dtype, s.nx) && !getenv("OF3D_NO_FUSED_T")) {
        const int rc = run_fast<T>(c, s, fp, in_dtype, nullptr, nullptr, t, own_lo, own_n, vx, vy, vz, rel, rel_f32);
        if (rc != kNotSupported) return rc;
    }
    const T* ic = ic_in; const T* dt0 = dt0_in;
    if (fp) {
        T* a = ic_out ? ic_out : ws_take<T>(c, s.n());
        T* b = dt0_out ? dt0_out : ws_take<T>(c, s.n());
        if (int rc = g.run_temporal(*fp, in_dtype, a, b)) return rc;
        ic = a; dt0 = b;
    }
    if (!spatial) return OF3D_OK;
    if (fast) {
        const int rc = run_fast<T>(c, s, nullptr, 0, ic, dt0, t, own_lo, own_n, vx, vy, vz, rel, rel_f32);   // marching kernels
        if (rc != kNotSupported) return rc;
    }
    // generic kernels: whole volume, then the owned planes are copied out
    const bool whole = own_lo == 0 && own_n == s.nz;
    const int64_t plane = s.ny * s.nx;
    T* o[4] = {vx, vy, vz, rel};
    if (!whole) for (int i = 0; i < 4; ++i) o[i] = (i == 2 && s.ndim == 2) ? nullptr : ws_take<T>(c, s.n());
    else if (rel_f32) o[3] = ws_take<T>(c, s.n());
    if (int rc = g.run_spatial(ic, dt0, o[0], o[1], o[2], o[3])) return rc;
    T* fin[4] = {vx, vy, vz, rel};
    for (int i = 0; i < 4; ++i) {
        if (!fin[i] || fin[i] == o[i]) continue;
        const T* srcp = o[i] + own_lo * plane;
        if (i == 3 && rel_f32) {
            if constexpr (sizeof(T) == 8) {
                StageScope span(c, OF3D_STAGE_GENERIC);           // epilogue of calc_flow.py:355-357
                narrow_f64_f32<<<grid_for(c, own_n * plane), 256, 0, c->stream>>>(srcp, reinterpret_cast<float*>(rel), own_n * plane);
                c->launches++;
            }
        } else {
            OF3D_CUDA_TRY(cudaMemcpyAsync(fin[i], srcp, (size_t)own_n * plane * sizeof(T), cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    return OF3D_OK;
}

template <typename T>
static int run_typed(of3d_ctx* c, const Shape& s, const FramePtrs* fp, int in_dtype, const of3d_taps* t, unsigned flags,
                     const void* ic_in, const void* dt0_in, void* ic_out, void* dt0_out, bool spatial, int64_t own_lo, int64_t own_n,
                     void* const dout[4], bool rel_f32) {
    if (flags & OF3D_FLAG_EXACT)
        return run_pipe<T, true>(c, s, fp, in_dtype, t, flags, (const T*)ic_in, (const T*)dt0_in, (T*)ic_out, (T*)dt0_out, spatial,
                                 own_lo, own_n, (T*)dout[0], (T*)dout[1], (T*)dout[2], (T*)dout[3], rel_f32);
    return run_pipe<T, false>(c, s, fp, in_dtype, t, flags, (const T*)ic_in, (const T*)dt0_in, (T*)ic_out, (T*)dt0_out, spatial,
                              own_lo, own_n, (T*)dout[0], (T*)dout[1], (T*)dout[2], (T*)dout[3], rel_f32);
}

// stage: 0 = whole operator from frames, 1 = temporal stage only (frames -> ic, dt0 device buffers),
//        2 = spatial stages from (ic, dt0) device buffers
// The inputs hold nz planes; in 3D the outputs hold the own_n planes [own_lo, own_lo + own_n) (own_n = 0: all of them).
static int flow_staged(of3d_ctx* c, int stage, int ndim, const void* const* frames, int in_dtype, int in_mem, int64_t nz,
                       int64_t ny, int64_t nx, const of3d_taps* t, int precision, unsigned flags, void* ic_dev, void* dt0_dev,
                       void* vx, void* vy, void* vz, void* rel, int out_mem, int64_t own_lo = 0, int64_t own_n = 0) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    if (ndim != 2 && ndim != 3) { set_error("ndim must be 2 or 3"); return OF3D_ERR_ARG; }
    if (int rc = check_taps(t)) return rc;
    // ndim == 2 with nz > 1: a BATCH of nz consecutive output timepoints of a 2D time-lapse (of3d_flow2d_batch): `frames`
    // holds nz + nT - 1 consecutive device frames and the outputs are (nz, ny, nx)
    const bool batch2d = ndim == 2 && nz > 1;
    if (nz < 1 || ny < 1 || nx < 1) { set_error("bad volume shape"); return OF3D_ERR_ARG; }
    if (batch2d && (stage != 0 || in_mem != OF3D_DEVICE || out_mem != OF3D_DEVICE || own_n != 0)) { set_error("2D batches take device frames and device outputs"); return OF3D_ERR_ARG; }
    const int n_frames = stage == 2 ? 0 : (batch2d ? (int)nz + t->nT - 1 : t->nT);
    if (n_frames > kMaxFrames) { set_error("too many frames in one 2D batch"); return OF3D_ERR_ARG; }
    if (own_n == 0) { own_lo = 0; own_n = nz; }
    if (own_lo < 0 || own_n < 1 || own_lo + own_n > nz) { set_error("bad z range"); return OF3D_ERR_ARG; }
    if (precision != OF3D_FP64 && precision != OF3D_FP32) { set_error("precision must be OF3D_FP64 or OF3D_FP32"); return OF3D_ERR_ARG; }
    if ((in_mem != OF3D_HOST && in_mem != OF3D_DEVICE) || (out_mem != OF3D_HOST && out_mem != OF3D_DEVICE)) { set_error("bad memory space"); return OF3D_ERR_ARG; }
    if (stage != 2) {
        if (!dtype_size(in_dtype)) { set_error("unsupported input dtype"); return OF3D_ERR_ARG; }
        if (!frames) { set_error("null image pointer"); return OF3D_ERR_ARG; }
        for (int k = 0; k < n_frames; ++k) if (!frames[k]) { set_error("null frame pointer"); return OF3D_ERR_ARG; }
    }
    if (stage != 0 && (!ic_dev || !dt0_dev)) { set_error("null ic/dt0 pointer"); return OF3D_ERR_ARG; }
    const bool rel_f32 = stage != 1 && precision == OF3D_FP64 && (flags & OF3D_FLAG_REL_F32);
    {
        // device buffers of the compute type are accessed with 8-byte (fp64) / 4-byte (fp32) vector-free loads and
        // cp.async: they must be naturally aligned
        const uintptr_t al = precision == OF3D_FP32 ? 4 : 8;
        const void* chk[6] = {stage != 0 ? ic_dev : nullptr, stage != 0 ? dt0_dev : nullptr,
                              out_mem == OF3D_DEVICE ? vx : nullptr, out_mem == OF3D_DEVICE ? vy : nullptr,
                              out_mem == OF3D_DEVICE ? vz : nullptr, out_mem == OF3D_DEVICE ? rel : nullptr};
        if (rel_f32 && reinterpret_cast<uintptr_t>(rel) % 4) { set_error("device buffers must be aligned to the compute type"); return OF3D_ERR_ARG; }
        if (rel_f32) chk[5] = nullptr;
        for (const void* q : chk)
            if (q && reinterpret_cast<uintptr_t>(q) % al) { set_error("device buffers must be aligned to the compute type"); return OF3D_ERR_ARG; }
        if (stage != 2 && in_mem == OF3D_DEVICE)
            for (int k = 0; k < n_frames; ++k)
                if (reinterpret_cast<uintptr_t>(frames[k]) % dtype_size(in_dtype)) { set_error("frame pointers must be aligned to the image dtype"); return OF3D_ERR_ARG; }
    }
    if (stage != 1 && (!vx || !vy || !rel || (ndim == 3 && !vz))) { set_error("null image or output pointer"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));

    const Shape s{ndim, nz, ny, nx};
    const int64_t n = s.n(), n_own = own_n * ny * nx;
    const size_t ts = precision == OF3D_FP32 ? 4 : 8;
    {
        const bool fast = !(flags & (OF3D_FLAG_EXACT | OF3D_FLAG_GENERIC)) && fast_supported(t);
        // (host frames are staged 512-byte aligned, device frames are checked again where the kernel is chosen)
        bool fused = fast && stage == 0 && ndim == 3 && !getenv("OF3D_NO_FUSED_T");
        if (fused) {
            FramePtrs probe;
            memset(&probe, 0, sizeof(probe));
            if (in_mem == OF3D_DEVICE) for (int k = 0; k < t->nT; ++k) probe.p[k] = frames[k];
            fused = fused_temporal_ok(t, probe, in_dtype, nx);
        }
        // (the temporal stage alone writes straight into the caller's ic / dt0: it only needs the staging of host frames)
        const size_t need = stage == 1 ? (in_mem == OF3D_HOST ? (size_t)t->nT * align_up((size_t)n * dtype_size(in_dtype)) + 8192 : 8192)
                                       : plan_bytes(ndim, stage == 2 ? 0 : t->nT, nz, own_n, ny * nx, t->nW, in_dtype, precision,
                                                    stage == 2 ? OF3D_DEVICE : in_mem, out_mem, fast ? 1 : 0, fused || stage == 2, rel_f32);
        if (int rc = ws_ensure(c, need)) return rc;
    }
    c->ws_off = 0;

    FramePtrs fp;
    memset(&fp, 0, sizeof(fp));
    if (stage != 2) {
        if (in_mem == OF3D_HOST) {
            const size_t fb = (size_t)n * dtype_size(in_dtype);
            for (int k = 0; k < t->nT; ++k) {
                char* d = ws_take<char>(c, fb);
                OF3D_CUDA_TRY(cudaMemcpyAsync(d, frames[k], fb, cudaMemcpyHostToDevice, c->stream));
                fp.p[k] = d;
            }
        } else {
            for (int k = 0; k < n_frames; ++k) fp.p[k] = frames[k];
        }
    }
    if (batch2d) {
        // temporal stage plane by plane (every output timepoint has its own window of frames), then the spatial stages
        // of all nz planes in ONE set of launches: whole waves instead of five launch tails per frame
        char* ic = ws_take<char>(c, (size_t)n * ts);
        char* dt0 = ws_take<char>(c, (size_t)n * ts);
        const Shape s1{2, 1, ny, nx};
        void* none[4] = {nullptr, nullptr, nullptr, nullptr};
        // all timepoints of the batch in one launch, every frame read once (temporal_vec_batch): 8/16-bit frames, the
        // reference's default seven temporal taps, 16-byte aligned frames and planes
        bool batched_t = !(flags & OF3D_FLAG_EXACT) && t->nT == 7 && (in_dtype == OF3D_U8 || in_dtype == OF3D_U16) &&
                         taps_symmetric(t->T, t->nT, -1.0) && t->T[t->nT / 2] == 0.0 && ((size_t)ny * nx * dtype_size(in_dtype)) % 16 == 0 &&
                         !getenv("OF3D_NO_BATCHED_T");
        for (int k = 0; batched_t && k < n_frames; ++k) batched_t = reinterpret_cast<uintptr_t>(fp.p[k]) % 16 == 0;
        if (batched_t) {
            const int64_t nvec = ny * nx * (int64_t)dtype_size(in_dtype) / 16;
            StageScope span(c, OF3D_STAGE_TEMPORAL);
            auto go = [&](auto tin, auto tt) {
                using Tin = decltype(tin); using TT = decltype(tt);
                TapsHalf<TT, 3> tw;
                for (int l = 0; l <= 3; ++l) tw.w[l] = (TT)t->T[3 + l];
                temporal_vec_batch<Tin, TT, 7><<<grid_for(c, nvec), 256, 0, c->stream>>>(fp, tw, reinterpret_cast<TT*>(ic), reinterpret_cast<TT*>(dt0),
                                                                                         nvec, (int)nz, ny * nx);
            };
            if (precision == OF3D_FP64) { if (in_dtype == OF3D_U8) go(uint8_t(), double()); else go(uint16_t(), double()); }
            else { if (in_dtype == OF3D_U8) go(uint8_t(), float()); else go(uint16_t(), float()); }
            c->launches++;
            OF3D_CUDA_TRY(cudaGetLastError());
        }
        for (int64_t j = 0; !batched_t && j < nz; ++j) {
            FramePtrs fj;
            memset(&fj, 0, sizeof(fj));
            for (int k = 0; k < t->nT; ++k) fj.p[k] = fp.p[j + k];
            void* icj = ic + (size_t)j * ny * nx * ts; void* dtj = dt0 + (size_t)j * ny * nx * ts;
            const int rc = precision == OF3D_FP64
                ? run_typed<double>(c, s1, &fj, in_dtype, t, flags, nullptr, nullptr, icj, dtj, false, 0, 1, none, false)
                : run_typed<float>(c, s1, &fj, in_dtype, t, flags, nullptr, nullptr, icj, dtj, false, 0, 1, none, false);
            if (rc) return rc;
        }
        void* dout2[4] = {vx, vy, nullptr, rel};
        const int rc = precision == OF3D_FP64
            ? run_typed<double>(c, s, nullptr, in_dtype, t, flags, ic, dt0, nullptr, nullptr, true, 0, nz, dout2, false)
            : run_typed<float>(c, s, nullptr, in_dtype, t, flags, ic, dt0, nullptr, nullptr, true, 0, nz, dout2, false);
        if (rc) return rc;
        OF3D_CUDA_TRY(cudaGetLastError());
        if (!c->async) OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
        return OF3D_OK;
    }
    void* out[4] = {vx, vy, vz, rel};
    void* dout[4] = {vx, vy, vz, rel};
    const int nout = ndim + 1;
    const int oidx3[4] = {0, 1, 2, 3}, oidx2[3] = {0, 1, 3};
    const int* oidx = ndim == 3 ? oidx3 : oidx2;
    if (stage != 1 && out_mem == OF3D_HOST)
        for (int i = 0; i < nout; ++i) dout[oidx[i]] = ws_take<char>(c, (size_t)n_own * ts);

    const FramePtrs* fpp = stage == 2 ? nullptr : &fp;
    int rc;
    if (precision == OF3D_FP64)
        rc = run_typed<double>(c, s, fpp, in_dtype, t, flags, ic_dev, dt0_dev, stage == 1 ? ic_dev : nullptr,
                               stage == 1 ? dt0_dev : nullptr, stage != 1, own_lo, own_n, dout, rel_f32);
    else
        rc = run_typed<float>(c, s, fpp, in_dtype, t, flags, ic_dev, dt0_dev, stage == 1 ? ic_dev : nullptr,
                              stage == 1 ? dt0_dev : nullptr, stage != 1, own_lo, own_n, dout, false);
    if (rc) return rc;
    OF3D_CUDA_TRY(cudaGetLastError());
    if (stage != 1 && out_mem == OF3D_HOST)
        for (int i = 0; i < nout; ++i)
            OF3D_CUDA_TRY(cudaMemcpyAsync(out[oidx[i]], dout[oidx[i]], (size_t)n_own * ((rel_f32 && oidx[i] == 3) ? 4 : ts),
                                          cudaMemcpyDeviceToHost, c->stream));
    const bool all_device = (stage == 2 || in_mem == OF3D_DEVICE) && (stage == 1 || out_mem == OF3D_DEVICE);
    if (!(c->async && all_device)) OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OF3D_OK;
}

static int flow_frames_impl(of3d_ctx* c, int ndim, const void* const* frames, int in_dtype, int in_mem, int64_t nz,
                            int64_t ny, int64_t nx, const of3d_taps* t, int precision, unsigned flags, void* vx, void* vy,
                            void* vz, void* rel, int out_mem) {
    return flow_staged(c, 0, ndim, frames, in_dtype, in_mem, nz, ny, nx, t, precision, flags, nullptr, nullptr, vx, vy, vz, rel, out_mem);
}

static int flow_contig(of3d_ctx* c, int ndim, const void* images, int in_dtype, int in_mem, int64_t nt, int64_t nz, int64_t ny,
                       int64_t nx, const of3d_taps* t, int precision, unsigned flags, void* vx, void* vy, void* vz, void* rel,
                       int out_mem) {
    if (int rc = check_taps(t)) return rc;
    if (!images) { set_error("images is null"); return OF3D_ERR_ARG; }
    if (nt < t->nT || !(nt & 1)) { set_error("nt must be odd and >= the number of temporal taps"); return OF3D_ERR_ARG; }
    if (!dtype_size(in_dtype)) { set_error("unsupported input dtype"); return OF3D_ERR_ARG; }
    const int64_t c0 = (nt + 1) / 2 - 1 - t->nT / 2;  // first frame the centre slice of the t-filter touches
    const size_t fb = (size_t)nz * ny * nx * dtype_size(in_dtype);
    const void* frames[kMaxFrames];
    for (int k = 0; k < t->nT; ++k) frames[k] = (const char*)images + (size_t)(c0 + k) * fb;
    return flow_frames_impl(c, ndim, frames, in_dtype, in_mem, nz, ny, nx, t, precision, flags, vx, vy, vz, rel, out_mem);
}


// k-th smallest (0-based) non-NaN element by MSB-first radix select; *n_nan = number of NaNs in the array
template <typename Tv, typename Key>
static int order_stat(of3d_ctx* c, const Tv* x, int64_t n, int64_t k, double* out, int64_t* n_nan) {
    constexpr int KEYBITS = sizeof(Key) * 8;
    unsigned long long* d = ws_take<unsigned long long>(c, kRadixBins + 1);
    std::vector<unsigned long long> h(kRadixBins + 1);
    Key prefix = 0, mask = 0;
    int shift = KEYBITS;
    while (shift > 0) {
        const int bits = std::min(kRadixBits, shift);
        shift -= bits;
        OF3D_CUDA_TRY(cudaMemsetAsync(d, 0, (kRadixBins + 1) * sizeof(unsigned long long), c->stream));
        radix_hist<Tv, Key><<<grid_for(c, n), 256, 0, c->stream>>>(x, n, prefix, mask, shift, (1u << bits) - 1u, d, d + kRadixBins);
        c->launches++;
        OF3D_CUDA_TRY(cudaMemcpyAsync(h.data(), d, (kRadixBins + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (n_nan) *n_nan = (int64_t)h[kRadixBins];
        const int nb = 1 << bits;
        int b = 0;
        for (; b < nb; ++b) {
            if ((unsigned long long)k < h[b]) break;
            k -= (int64_t)h[b];
        }
        if (b == nb) { set_error("rank beyond the number of non-NaN elements"); return OF3D_ERR_ARG; }
        prefix |= (Key)b << shift;
        mask |= (Key)(nb - 1) << shift;
    }
    // invert sort_key
    const Key top = (Key)1 << (KEYBITS - 1);
    const Key bits = (prefix & top) ? (prefix & ~top) : ~prefix;
    Tv v;
    memcpy(&v, &bits, sizeof(v));
    *out = (double)v;
    return OF3D_OK;
}

}  // namespace of3d

// =============================================================================================
extern "C" {

OF3D_API int of3d_version(void) { return OF3D_VERSION; }
OF3D_API const char* of3d_last_error(void) { return of3d::g_err.c_str(); }

OF3D_API int of3d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

OF3D_API int of3d_create(int device, of3d_ctx** out) {
    if (!out) { set_error("out is null"); return OF3D_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (libof3d has no CPU fallback)");
        return OF3D_ERR_NODEVICE;
    }
    if (device < 0 || device >= n) { set_error("device index out of range"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(device));
    of3d_ctx* c = new (std::nothrow) of3d_ctx();
    if (!c) { set_error("out of host memory"); return OF3D_ERR_NOMEM; }
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; set_error(std::string("cudaStreamCreate failed: ") + cudaGetErrorString(e)); return OF3D_ERR_CUDA; }
    c->stream = c->own_stream;
    *out = c;
    return OF3D_OK;
}

OF3D_API int of3d_destroy(of3d_ctx* c) {
    if (!c) return OF3D_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) of3d_comm_destroy(c);
    if (c->own_stream) { cudaStreamSynchronize(c->own_stream); cudaStreamDestroy(c->own_stream); }
    if (c->s_comm) cudaStreamDestroy(c->s_comm);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    for (auto& sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->ws) cudaFree(c->ws);
    if (c->win) cudaFree(c->win);
    if (c->pipe) cudaFree(c->pipe);
    if (c->s_up) cudaStreamDestroy(c->s_up);
    if (c->s_dn) cudaStreamDestroy(c->s_dn);
    if (c->ev_up) cudaEventDestroy(c->ev_up);
    if (c->ev_c) cudaEventDestroy(c->ev_c);
    for (auto e : c->ev_dn) if (e) cudaEventDestroy(e);
    for (auto& pt : c->parts) cudaEventDestroy(pt.ev);
    for (auto e : c->part_pool) cudaEventDestroy(e);
    delete c;
    return OF3D_OK;
}

OF3D_API size_t of3d_workspace_bytes(int ndim, int64_t nt_taps, int64_t nz, int64_t ny, int64_t nx, int in_dtype, int precision,
                            int in_mem, int out_mem) {
    if (ndim != 2 && ndim != 3) return 0;
    return plan_bytes(ndim, nt_taps, nz, nz, ny * nx, 0, in_dtype, precision, in_mem, out_mem, -1, false, true);
}

OF3D_API int of3d_reserve(of3d_ctx* c, size_t bytes) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    return ws_ensure(c, bytes);
}

OF3D_API int of3d_flow3d(of3d_ctx* ctx, const void* images, int in_dtype, int in_mem, int64_t nt, int64_t nz, int64_t ny, int64_t nx,
                const of3d_taps* taps, int precision, unsigned flags, void* vx, void* vy, void* vz, void* rel, int out_mem) {
    return flow_contig(ctx, 3, images, in_dtype, in_mem, nt, nz, ny, nx, taps, precision, flags, vx, vy, vz, rel, out_mem);
}

OF3D_API int of3d_flow2d(of3d_ctx* ctx, const void* images, int in_dtype, int in_mem, int64_t nt, int64_t ny, int64_t nx,
                const of3d_taps* taps, int precision, unsigned flags, void* vx, void* vy, void* rel, int out_mem) {
    return flow_contig(ctx, 2, images, in_dtype, in_mem, nt, 1, ny, nx, taps, precision, flags, vx, vy, nullptr, rel, out_mem);
}

OF3D_API int of3d_flow2d_batch(of3d_ctx* ctx, const void* const* frames, int in_dtype, int64_t n_out, int64_t ny, int64_t nx,
                               const of3d_taps* taps, int precision, unsigned flags, void* vx, void* vy, void* rel) {
    return flow_staged(ctx, 0, 2, frames, in_dtype, OF3D_DEVICE, n_out, ny, nx, taps, precision, flags, nullptr, nullptr, vx, vy, nullptr, rel, OF3D_DEVICE);
}

OF3D_API int of3d_flow_frames(of3d_ctx* ctx, int ndim, const void* const* frames, int in_dtype, int in_mem, int64_t nz, int64_t ny,
                     int64_t nx, const of3d_taps* taps, int precision, unsigned flags, void* vx, void* vy, void* vz, void* rel,
                     int out_mem) {
    return flow_frames_impl(ctx, ndim, frames, in_dtype, in_mem, nz, ny, nx, taps, precision, flags, vx, vy, vz, rel, out_mem);
}

OF3D_API int64_t of3d_window_slab(int ndim, int64_t nz, int64_t ny, int64_t nx, const of3d_taps* t) {
    if (ndim != 3 || !t || getenv("OF3D_NO_SLAB_PIPELINE")) return 0;
    const int64_t H = std::max(std::max(t->nD, t->nG), t->nS) / 2 + t->nW / 2;
    const int64_t slab = 16;
    // worth it when the copies dominate (>= 32 Mvoxel) and the extension (slab + 2H) stays well below the volume
    if ((nz * ny * nx < (int64_t(32) << 20) && !getenv("OF3D_FORCE_SLAB_PIPELINE")) || nz < 4 * slab || nz < 4 * H) return 0;
    return slab;
}

OF3D_API int of3d_window_upload(of3d_ctx* c, int k, int n_frames, const void* host_src, size_t frame_bytes, size_t offset, size_t bytes) {
    if (!c || !host_src || n_frames < 1 || n_frames > kMaxFrames || k < 0 || k >= n_frames || frame_bytes == 0 || bytes == 0 ||
        offset + bytes > frame_bytes) { set_error("bad argument"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    if (!c->s_up) {
        OF3D_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking));
        OF3D_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_dn, cudaStreamNonBlocking));
        OF3D_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
        OF3D_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_c, cudaEventDisableTiming));
        for (auto& e : c->ev_dn) OF3D_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const size_t slot = align_up(frame_bytes);
    const bool first = k == 0 && offset == 0;
    if (first) {
        // the previous window may still be read by kernels of an asynchronous call
        OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
        OF3D_CUDA_TRY(cudaStreamSynchronize(c->s_up));
        if (c->win_cap < slot * n_frames) {
            if (c->win) cudaFree(c->win);
            c->win = nullptr; c->win_cap = 0;
            cudaError_t e = cudaMalloc(&c->win, slot * n_frames);
            if (e != cudaSuccess) { cudaGetLastError(); set_error("device window allocation of " + std::to_string(slot * n_frames) + " bytes failed: out of memory"); return OF3D_ERR_NOMEM; }
            c->win_cap = slot * n_frames;
        }
        c->win_frame = frame_bytes; c->win_n = n_frames;
        for (auto& pt : c->parts) c->part_pool.push_back(pt.ev);
        c->parts.clear();
    } else if (frame_bytes != c->win_frame || n_frames != c->win_n) { set_error("a window starts with frame 0, offset 0 and keeps one frame size"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaMemcpyAsync(c->win + (size_t)k * slot + offset, host_src, bytes, cudaMemcpyHostToDevice, c->s_up));
    cudaEvent_t ev = nullptr;
    if (!c->part_pool.empty()) { ev = c->part_pool.back(); c->part_pool.pop_back(); }
    else OF3D_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    OF3D_CUDA_TRY(cudaEventRecord(ev, c->s_up));
    c->parts.push_back({offset, ev});
    OF3D_CUDA_TRY(cudaEventRecord(c->ev_up, c->s_up));
    return OF3D_OK;
}

// The synchronous host call in z slabs: while slab s is copied back, slab s + 1 is computed and the planes of later
// slabs are still arriving.  A slab is computed from its extension by H = R_gradient + R_window planes (the z support of
// the operator) with the owned-range pipeline, which returns the interior only: bit-identical to the whole-volume run.
static int window_flow_pipelined(of3d_ctx* c, int in_dtype, int64_t nz, int64_t ny, int64_t nx, const of3d_taps* t, int precision,
                                 unsigned flags, void* const hout[4], int64_t slab) {
    const int64_t plane = ny * nx;
    const size_t ts = precision == OF3D_FP32 ? 4 : 8, ib = dtype_size(in_dtype);
    const int64_t H = std::max(std::max(t->nD, t->nG), t->nS) / 2 + t->nW / 2;
    const int64_t ext_max = std::min(nz, slab + 2 * H);
    const size_t osz[4] = {ts, ts, ts, (precision == OF3D_FP64 && (flags & OF3D_FLAG_REL_F32)) ? (size_t)4 : ts};
    // device buffers: two sets of four slab outputs
    const size_t out_b = align_up((size_t)slab * plane * 8);
    const size_t need = 8 * out_b;
    if (c->pipe_cap < need) {
        OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
        OF3D_CUDA_TRY(cudaStreamSynchronize(c->s_dn));
        if (c->pipe) cudaFree(c->pipe);
        c->pipe = nullptr; c->pipe_cap = 0;
        cudaError_t e = cudaMalloc(&c->pipe, need);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("device allocation of " + std::to_string(need) + " bytes failed: out of memory"); return OF3D_ERR_NOMEM; }
        c->pipe_cap = need;
    }
    // one arena size for every stage call of the pipeline (no re-allocation in flight)
    {
        const bool fast = !(flags & (OF3D_FLAG_EXACT | OF3D_FLAG_GENERIC)) && fast_supported(t);
        if (int rc = ws_ensure(c, plan_bytes(3, t->nT, ext_max, std::min(slab, nz), plane, t->nW, in_dtype, precision, OF3D_DEVICE, OF3D_DEVICE,
                                             fast ? 1 : 0, false, (flags & OF3D_FLAG_REL_F32) != 0))) return rc;
    }
    char* ext[2][4];
    for (int q = 0; q < 2; ++q) for (int i = 0; i < 4; ++i) ext[q][i] = c->pipe + (size_t)(q * 4 + i) * out_b;
    const size_t slot = align_up(c->win_frame);
    const int saved_async = c->async;
    c->async = 1;                                               // the stage calls below must not synchronise
    // (a lambda so that every early error return still restores the context and drains the streams below)
    auto run = [&]() -> int {
        int64_t tz = 0;                                         // planes whose arrival the compute stream already waits for
        int ns = 0;
        for (int64_t a = 0; a < nz; a += slab, ++ns) {
            const int64_t b = std::min(nz, a + slab), a2 = std::max<int64_t>(0, a - H), b2 = std::min(nz, b + H);
            if (tz < b2) {
                // wait (on the device) for every uploaded piece that holds planes below b2
                const size_t limit = (size_t)b2 * plane * ib;
                cudaEvent_t ev = nullptr;
                for (const auto& pt : c->parts) if (pt.off < limit) ev = pt.ev;
                if (ev) OF3D_CUDA_TRY(cudaStreamWaitEvent(c->stream, ev, 0));
                tz = b2;
            }
            const int q = ns & 1;
            if (ns >= 2) OF3D_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_dn[q], 0));   // the copy-back of slab ns - 2 has left the buffers
            const void* frames[kMaxFrames];
            for (int k = 0; k < t->nT; ++k) frames[k] = c->win + (size_t)k * slot + (size_t)a2 * plane * ib;
            if (int rc = flow_staged(c, 0, 3, frames, in_dtype, OF3D_DEVICE, b2 - a2, ny, nx, t, precision, flags, nullptr, nullptr,
                                     ext[q][0], ext[q][1], ext[q][2], ext[q][3], OF3D_DEVICE, a - a2, b - a)) return rc;
            OF3D_CUDA_TRY(cudaEventRecord(c->ev_c, c->stream));
            OF3D_CUDA_TRY(cudaStreamWaitEvent(c->s_dn, c->ev_c, 0));
            for (int i = 0; i < 4; ++i)
                OF3D_CUDA_TRY(cudaMemcpyAsync((char*)hout[i] + (size_t)a * plane * osz[i], ext[q][i], (size_t)(b - a) * plane * osz[i],
                                              cudaMemcpyDeviceToHost, c->s_dn));
            OF3D_CUDA_TRY(cudaEventRecord(c->ev_dn[q], c->s_dn));
        }
        return OF3D_OK;
    };
    const int rc = run();
    c->async = saved_async;
    cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->s_dn);
    if (rc) return rc;
    OF3D_CUDA_TRY(e1);
    OF3D_CUDA_TRY(e2);
    return OF3D_OK;
}

OF3D_API int of3d_window_flow(of3d_ctx* c, int ndim, int in_dtype, int64_t nz, int64_t ny, int64_t nx, const of3d_taps* taps, int precision,
                              unsigned flags, void* vx, void* vy, void* vz, void* rel, int out_mem) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    if (int rc = check_taps(taps)) return rc;
    if (!c->win || c->win_n != taps->nT || c->win_frame != (size_t)(nz * ny * nx) * dtype_size(in_dtype)) {
        set_error("of3d_window_flow: the uploaded window does not match (frames or frame size)");
        return OF3D_ERR_ARG;
    }
    if (precision != OF3D_FP64 && precision != OF3D_FP32) { set_error("precision must be OF3D_FP64 or OF3D_FP32"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    // large 3D volumes going back to the host: z-slab pipeline (upload, compute and copy-back overlap)
    const int64_t slab = of3d_window_slab(ndim, nz, ny, nx, taps);
    if (out_mem == OF3D_HOST && slab > 0 && vx && vy && vz && rel) {
        void* hout[4] = {vx, vy, vz, rel};
        return window_flow_pipelined(c, in_dtype, nz, ny, nx, taps, precision, flags, hout, slab);
    }
    OF3D_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_up, 0));
    const void* frames[kMaxFrames];
    const size_t slot = align_up(c->win_frame);
    for (int k = 0; k < taps->nT; ++k) frames[k] = c->win + (size_t)k * slot;
    return flow_frames_impl(c, ndim, frames, in_dtype, OF3D_DEVICE, nz, ny, nx, taps, precision, flags, vx, vy, vz, rel, out_mem);
}

OF3D_API int of3d_temporal(of3d_ctx* ctx, int ndim, const void* const* frames, int in_dtype, int in_mem, int64_t nz, int64_t ny,
                           int64_t nx, const of3d_taps* taps, int precision, unsigned flags, void* ic_dev, void* dt0_dev) {
    return flow_staged(ctx, 1, ndim, frames, in_dtype, in_mem, nz, ny, nx, taps, precision, flags, ic_dev, dt0_dev, nullptr, nullptr,
                       nullptr, nullptr, OF3D_DEVICE);
}

OF3D_API int of3d_flow_from_dt(of3d_ctx* ctx, int ndim, const void* ic_dev, const void* dt0_dev, int64_t nz, int64_t ny, int64_t nx,
                               const of3d_taps* taps, int precision, unsigned flags, void* vx, void* vy, void* vz, void* rel,
                               int out_mem) {
    return flow_staged(ctx, 2, ndim, nullptr, OF3D_U16, OF3D_DEVICE, nz, ny, nx, taps, precision, flags, const_cast<void*>(ic_dev),
                       const_cast<void*>(dt0_dev), vx, vy, vz, rel, out_mem);
}

OF3D_API void* of3d_stream(of3d_ctx* c) { return c ? (void*)c->stream : nullptr; }

OF3D_API int of3d_set_stream(of3d_ctx* c, void* cuda_stream) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));            // nothing of the context is left behind on the old stream
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return OF3D_OK;
}

OF3D_API int of3d_set_async(of3d_ctx* c, int enable) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    c->async = enable ? 1 : 0;
    return OF3D_OK;
}

OF3D_API int of3d_sync(of3d_ctx* c) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OF3D_OK;
}

OF3D_API int64_t of3d_launch_count(of3d_ctx* c) { return c ? c->launches : 0; }

OF3D_API int of3d_set_profile(of3d_ctx* c, int enable) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    c->profile = enable ? 1 : 0;
    return OF3D_OK;
}

OF3D_API int of3d_stage_times(of3d_ctx* c, double* ms, int64_t* launches) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (const auto& sp : c->spans) {
        float t = 0.f;
        OF3D_CUDA_TRY(cudaEventElapsedTime(&t, sp.a, sp.b));
        c->stage_ms[sp.stage] += t;
        c->stage_launches[sp.stage]++;
        c->ev_pool.push_back(sp.a);
        c->ev_pool.push_back(sp.b);
    }
    c->spans.clear();
    for (int i = 0; i < OF3D_N_STAGES; ++i) {
        if (ms) ms[i] = c->stage_ms[i];
        if (launches) launches[i] = c->stage_launches[i];
        c->stage_ms[i] = 0.0;
        c->stage_launches[i] = 0;
    }
    return OF3D_OK;
}

OF3D_API const char* of3d_stage_name(int stage) {
    static const char* names[OF3D_N_STAGES] = {"temporal", "gradient_xy", "gradient_z", "products_window_z", "window_xy_solve", "generic"};
    return (stage >= 0 && stage < OF3D_N_STAGES) ? names[stage] : "";
}

OF3D_API int of3d_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) { set_error("ptr is null"); return OF3D_ERR_ARG; }
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); set_error(std::string("cudaHostAlloc failed: ") + cudaGetErrorString(e)); return OF3D_ERR_NOMEM; }
    return OF3D_OK;
}

OF3D_API int of3d_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
    return OF3D_OK;
}

OF3D_API int of3d_order_stats(of3d_ctx* c, const void* data_dev, int is_f64, int64_t n, int64_t k_lo, int64_t k_hi, double* out_lo,
                              double* out_hi, int64_t* n_nan) {
    if (!c || !data_dev || n < 1 || k_lo < 0 || k_hi < k_lo || k_hi >= n || !out_lo || !out_hi) { set_error("bad argument"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    if (int rc = ws_ensure(c, 1 << 20)) return rc;
    c->ws_off = 0;
    int64_t nn = 0;
    int rc = is_f64 ? order_stat<double, uint64_t>(c, (const double*)data_dev, n, k_lo, out_lo, &nn)
                    : order_stat<float, uint32_t>(c, (const float*)data_dev, n, k_lo, out_lo, &nn);
    if (n_nan) *n_nan = nn;
    if (rc || nn > 0) { if (nn > 0) *out_hi = *out_lo; return rc; }      // NaNs present: the caller returns NaN like NumPy
    if (k_hi == k_lo) { *out_hi = *out_lo; return OF3D_OK; }
    return is_f64 ? order_stat<double, uint64_t>(c, (const double*)data_dev, n, k_hi, out_hi, nullptr)
                  : order_stat<float, uint32_t>(c, (const float*)data_dev, n, k_hi, out_hi, nullptr);
}

OF3D_API int of3d_mask_derive(of3d_ctx* c, const void* vx, const void* vy, const void* vz, const void* rel, int v_f64, int rel_f64,
                              int64_t n, double thresh, double xyscale, double zscale, double tscale, void* ox, void* oy, void* oz,
                              void* mag, void* theta, void* phi) {
    if (!c || !vx || !vy || !rel || !ox || !oy || !mag || !theta || n < 1 || (vz && (!oz || !phi))) { set_error("bad argument"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    const int grid = grid_for(c, n);
#define OF3D_MD(TV, TR)                                                                                                        \
    mask_derive<TV, TR><<<grid, 256, 0, c->stream>>>((const TV*)vx, (const TV*)vy, (const TV*)vz, (const TR*)rel, n, (TR)thresh,  \
                                                    (TV)xyscale, (TV)zscale, (TV)tscale, (TV*)ox, (TV*)oy, (TV*)oz, (TV*)mag,    \
                                                    (TV*)theta, (TV*)phi)
    if (v_f64 && rel_f64) OF3D_MD(double, double);
    else if (v_f64) OF3D_MD(double, float);
    else if (rel_f64) OF3D_MD(float, double);
    else OF3D_MD(float, float);
#undef OF3D_MD
    c->launches++;
    OF3D_CUDA_TRY(cudaGetLastError());
    OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OF3D_OK;
}

// ---------------------------------------------------------------------------------------------
// z-slab sharding.  NCCL is bound at run time (dlopen): a process that never shards needs no NCCL, and a
// process that already holds one (PyTorch) shares it.
namespace {
struct NcclId { char b[128]; };                                  // ncclUniqueId
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.h) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
        a.Send = (decltype(a.Send))dlsym(a.h, "ncclSend");
        a.Recv = (decltype(a.Recv))dlsym(a.h, "ncclRecv");
        a.GroupStart = (decltype(a.GroupStart))dlsym(a.h, "ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.h, "ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}
int nccl_fail(const char* what, int r) {
    NcclApi& n = nccl();
    set_error(std::string(what) + " failed: " + ((n.GetErrorString && r) ? n.GetErrorString(r) : "NCCL is not available (libnccl.so.2 not found)"));
    return OF3D_ERR_CUDA;
}
constexpr int kNcclUint8 = 1;                                    // ncclUint8
}  // namespace

OF3D_API int of3d_comm_unique_id(void* id128) {
    if (!id128) { set_error("id is null"); return OF3D_ERR_ARG; }
    NcclApi& n = nccl();
    if (!n.ok) return nccl_fail("loading NCCL", 0);
    NcclId id;
    if (int r = n.GetUniqueId(&id)) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id128, &id, sizeof(id));
    return OF3D_OK;
}

OF3D_API int of3d_comm_init(of3d_ctx* c, const void* id128, int nranks, int rank) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad argument"); return OF3D_ERR_ARG; }
    NcclApi& n = nccl();
    if (!n.ok) return nccl_fail("loading NCCL", 0);
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    if (c->comm) of3d_comm_destroy(c);
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    if (int r = n.CommInitRank(&c->comm, nranks, id, rank)) { c->comm = nullptr; return nccl_fail("ncclCommInitRank", r); }
    c->comm_rank = rank; c->comm_size = nranks;
    if (!c->s_comm) {
        OF3D_CUDA_TRY(cudaStreamCreateWithFlags(&c->s_comm, cudaStreamNonBlocking));
        OF3D_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
        OF3D_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    }
    return OF3D_OK;
}

OF3D_API int of3d_comm_destroy(of3d_ctx* c) {
    if (!c || !c->comm) return OF3D_OK;
    cudaSetDevice(c->device);
    if (c->s_comm) cudaStreamSynchronize(c->s_comm);
    nccl().CommDestroy(c->comm);
    c->comm = nullptr; c->comm_size = 1; c->comm_rank = 0; c->halo_pending = false;
    return OF3D_OK;
}

OF3D_API int of3d_halo_exchange(of3d_ctx* c, void* const* frames_ext, int n_frames, size_t plane_bytes, int64_t lo, int64_t own,
                                int64_t hi, int64_t send_dn, int64_t send_up) {
    if (!c || !frames_ext || n_frames < 1 || plane_bytes == 0 || lo < 0 || hi < 0 || own < 1 || send_dn < 0 || send_up < 0 ||
        send_dn > own || send_up > own) { set_error("bad argument"); return OF3D_ERR_ARG; }
    if (!c->comm) { set_error("of3d_halo_exchange: no communicator (of3d_comm_init)"); return OF3D_ERR_ARG; }
    for (int k = 0; k < n_frames; ++k) if (!frames_ext[k]) { set_error("null frame pointer"); return OF3D_ERR_ARG; }
    NcclApi& n = nccl();
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    const int r = c->comm_rank, last = c->comm_size - 1;
    // the owned planes may still be written by work queued on the compute stream
    OF3D_CUDA_TRY(cudaEventRecord(c->ev_ready, c->stream));
    OF3D_CUDA_TRY(cudaStreamWaitEvent(c->s_comm, c->ev_ready, 0));
    if (int e = n.GroupStart()) return nccl_fail("ncclGroupStart", e);
    int err = 0;
    for (int k = 0; k < n_frames && !err; ++k) {
        char* f = static_cast<char*>(frames_ext[k]);
        if (r > 0) {
            if (send_dn && !err) err = n.Send(f + (size_t)lo * plane_bytes, (size_t)send_dn * plane_bytes, kNcclUint8, r - 1, c->comm, c->s_comm);
            if (lo && !err) err = n.Recv(f, (size_t)lo * plane_bytes, kNcclUint8, r - 1, c->comm, c->s_comm);
        }
        if (r < last) {
            if (send_up && !err) err = n.Send(f + (size_t)(lo + own - send_up) * plane_bytes, (size_t)send_up * plane_bytes, kNcclUint8, r + 1, c->comm, c->s_comm);
            if (hi && !err) err = n.Recv(f + (size_t)(lo + own) * plane_bytes, (size_t)hi * plane_bytes, kNcclUint8, r + 1, c->comm, c->s_comm);
        }
    }
    const int e2 = n.GroupEnd();
    if (err) return nccl_fail("ncclSend/ncclRecv", err);
    if (e2) return nccl_fail("ncclGroupEnd", e2);
    OF3D_CUDA_TRY(cudaEventRecord(c->ev_halo, c->s_comm));
    c->halo_pending = true;
    return OF3D_OK;
}

OF3D_API int of3d_halo_exchange_centre(of3d_ctx* c, const void* centre_own, int in_dtype, void* stage, void* ic_ext, void* dt0_ext,
                                       int precision, int64_t plane_elems, int64_t lo, int64_t own, int64_t hi, int64_t send_dn,
                                       int64_t send_up) {
    if (!c || !centre_own || !stage || !ic_ext || !dt0_ext || plane_elems < 1 || lo < 0 || hi < 0 || own < 1 || send_dn < 0 || send_up < 0 ||
        send_dn > own || send_up > own) { set_error("bad argument"); return OF3D_ERR_ARG; }
    if (in_dtype != OF3D_U8 && in_dtype != OF3D_U16 && in_dtype != OF3D_I16) { set_error("of3d_halo_exchange_centre takes 8/16-bit integer frames"); return OF3D_ERR_ARG; }
    if (precision != OF3D_FP64 && precision != OF3D_FP32) { set_error("precision must be OF3D_FP64 or OF3D_FP32"); return OF3D_ERR_ARG; }
    if (!c->comm) { set_error("of3d_halo_exchange_centre: no communicator (of3d_comm_init)"); return OF3D_ERR_ARG; }
    NcclApi& n = nccl();
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    const int r = c->comm_rank, last = c->comm_size - 1;
    const size_t ib = dtype_size(in_dtype), ts = precision == OF3D_FP32 ? 4 : 8;
    const size_t pr = (size_t)plane_elems * ib, pt = (size_t)plane_elems * ts;      // plane bytes: raw, compute type
    // the owned planes may still be written by work queued on the compute stream
    OF3D_CUDA_TRY(cudaEventRecord(c->ev_ready, c->stream));
    OF3D_CUDA_TRY(cudaStreamWaitEvent(c->s_comm, c->ev_ready, 0));
    const char* cen = static_cast<const char*>(centre_own);
    char* st = static_cast<char*>(stage);                                          // [lo planes from below][hi planes from above]
    char* dt = static_cast<char*>(dt0_ext);
    if (int e = n.GroupStart()) return nccl_fail("ncclGroupStart", e);
    int err = 0;
    if (r > 0) {
        if (send_dn && !err) err = n.Send(cen, (size_t)send_dn * pr, kNcclUint8, r - 1, c->comm, c->s_comm);
        if (send_dn && !err) err = n.Send(dt + (size_t)lo * pt, (size_t)send_dn * pt, kNcclUint8, r - 1, c->comm, c->s_comm);
        if (lo && !err) err = n.Recv(st, (size_t)lo * pr, kNcclUint8, r - 1, c->comm, c->s_comm);
        if (lo && !err) err = n.Recv(dt, (size_t)lo * pt, kNcclUint8, r - 1, c->comm, c->s_comm);
    }
    if (r < last) {
        if (send_up && !err) err = n.Send(cen + (size_t)(own - send_up) * pr, (size_t)send_up * pr, kNcclUint8, r + 1, c->comm, c->s_comm);
        if (send_up && !err) err = n.Send(dt + (size_t)(lo + own - send_up) * pt, (size_t)send_up * pt, kNcclUint8, r + 1, c->comm, c->s_comm);
        if (hi && !err) err = n.Recv(st + (size_t)lo * pr, (size_t)hi * pr, kNcclUint8, r + 1, c->comm, c->s_comm);
        if (hi && !err) err = n.Recv(dt + (size_t)(lo + own) * pt, (size_t)hi * pt, kNcclUint8, r + 1, c->comm, c->s_comm);
    }
    const int e2 = n.GroupEnd();
    if (err) return nccl_fail("ncclSend/ncclRecv", err);
    if (e2) return nccl_fail("ncclGroupEnd", e2);
    // widen the raw planes that arrived into the halo planes of ic, on the exchange stream
    auto widen = [&](const char* src, char* dst, int64_t planes) {
        if (planes <= 0) return;
        const int64_t cnt = planes * plane_elems;
        const int grid = grid_for(c, cnt);
#define OF3D_WIDEN(TIN) \
        if (precision == OF3D_FP64) widen_planes<TIN, double><<<grid, 256, 0, c->s_comm>>>(reinterpret_cast<const TIN*>(src), reinterpret_cast<double*>(dst), cnt); \
        else widen_planes<TIN, float><<<grid, 256, 0, c->s_comm>>>(reinterpret_cast<const TIN*>(src), reinterpret_cast<float*>(dst), cnt);
        if (in_dtype == OF3D_U8) { OF3D_WIDEN(uint8_t) } else if (in_dtype == OF3D_U16) { OF3D_WIDEN(uint16_t) } else { OF3D_WIDEN(int16_t) }
#undef OF3D_WIDEN
        c->launches++;
    };
    char* ic = static_cast<char*>(ic_ext);
    if (r > 0) widen(st, ic, lo);
    if (r < last) widen(st + (size_t)lo * pr, ic + (size_t)(lo + own) * pt, hi);
    OF3D_CUDA_TRY(cudaGetLastError());
    OF3D_CUDA_TRY(cudaEventRecord(c->ev_halo, c->s_comm));
    c->halo_pending = true;
    return OF3D_OK;
}

// The owned range in chunks (interior chunks first: they overlap a halo exchange in flight); frames_ext != nullptr: raw
// frames, else (ic_ext, dt0_ext) volumes of the compute type
static int slab_run(of3d_ctx* c, const void* const* frames_ext, const void* ic_ext, const void* dt0_ext, int in_dtype, int64_t nz_ext,
                    int64_t ny, int64_t nx, int64_t own_lo, int64_t own_n, int64_t chunk_planes, const of3d_taps* t, int precision,
                    unsigned flags, void* vx, void* vy, void* vz, void* rel) {
    if (!c) { set_error("context is null"); return OF3D_ERR_ARG; }
    if (int rc = check_taps(t)) return rc;
    if (nz_ext < 1 || own_lo < 0 || own_n < 1 || own_lo + own_n > nz_ext || chunk_planes < 0) { set_error("bad z range"); return OF3D_ERR_ARG; }
    if ((!frames_ext && (!ic_ext || !dt0_ext)) || !vx || !vy || !vz || !rel) { set_error("null image or output pointer"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    const int64_t plane = ny * nx;
    const int64_t H = std::max(std::max(t->nD, t->nG), t->nS) / 2 + t->nW / 2;
    const size_t ts = precision == OF3D_FP32 ? 4 : 8, ib = frames_ext ? dtype_size(in_dtype) : ts;
    const size_t rs = (precision == OF3D_FP64 && (flags & OF3D_FLAG_REL_F32)) ? 4 : ts;
    if (!ib) { set_error("unsupported input dtype"); return OF3D_ERR_ARG; }
    const int64_t chunk = chunk_planes ? std::min(chunk_planes, own_n) : own_n;
    std::vector<int64_t> starts, late;
    for (int64_t a = own_lo; a < own_lo + own_n; a += chunk) {
        const int64_t b = std::min(own_lo + own_n, a + chunk);
        // a chunk waits for the exchange only if it reaches into a halo that exists (none at the ends of the volume)
        const bool need_lo = own_lo > 0 && a - H < own_lo, need_hi = own_lo + own_n < nz_ext && b + H > own_lo + own_n;
        const bool interior = !need_lo && !need_hi;
        (c->halo_pending && !interior ? late : starts).push_back(a);
    }
    const size_t n_first = starts.size();
    starts.insert(starts.end(), late.begin(), late.end());
    // one arena size for every chunk (no re-allocation while kernels are queued)
    {
        const int64_t ext_max = std::min(nz_ext, chunk + 2 * H);
        const bool fast = !(flags & (OF3D_FLAG_EXACT | OF3D_FLAG_GENERIC)) && fast_supported(t);
        if (int rc = ws_ensure(c, plan_bytes(3, frames_ext ? t->nT : 0, ext_max, chunk, plane, t->nW, in_dtype, precision, OF3D_DEVICE, OF3D_DEVICE,
                                             fast ? 1 : 0, false, (flags & OF3D_FLAG_REL_F32) != 0))) return rc;
    }
    const int saved_async = c->async;
    c->async = 1;
    int rc = OF3D_OK;
    for (size_t i = 0; i < starts.size() && !rc; ++i) {
        if (i == n_first && c->halo_pending) {
            cudaError_t e = cudaStreamWaitEvent(c->stream, c->ev_halo, 0);
            if (e != cudaSuccess) { set_error(std::string("cudaStreamWaitEvent failed: ") + cudaGetErrorString(e)); rc = OF3D_ERR_CUDA; break; }
            c->halo_pending = false;
        }
        const int64_t a = starts[i], b = std::min(own_lo + own_n, a + chunk);
        const int64_t a2 = std::max<int64_t>(0, a - H), b2 = std::min(nz_ext, b + H);
        const size_t o = (size_t)(a - own_lo) * plane;
        char* ov[4] = {(char*)vx + o * ts, (char*)vy + o * ts, (char*)vz + o * ts, (char*)rel + o * rs};
        if (frames_ext) {
            const void* frames[kMaxFrames];
            for (int k = 0; k < t->nT; ++k) {
                if (!frames_ext[k]) { set_error("null frame pointer"); rc = OF3D_ERR_ARG; break; }
                frames[k] = static_cast<const char*>(frames_ext[k]) + (size_t)a2 * plane * ib;
            }
            if (rc) break;
            rc = flow_staged(c, 0, 3, frames, in_dtype, OF3D_DEVICE, b2 - a2, ny, nx, t, precision, flags, nullptr, nullptr,
                             ov[0], ov[1], ov[2], ov[3], OF3D_DEVICE, a - a2, b - a);
        } else {
            rc = flow_staged(c, 2, 3, nullptr, OF3D_U16, OF3D_DEVICE, b2 - a2, ny, nx, t, precision, flags,
                             (char*)ic_ext + (size_t)a2 * plane * ts, (char*)dt0_ext + (size_t)a2 * plane * ts,
                             ov[0], ov[1], ov[2], ov[3], OF3D_DEVICE, a - a2, b - a);
        }
    }
    if (c->halo_pending) { cudaStreamWaitEvent(c->stream, c->ev_halo, 0); c->halo_pending = false; }   // later calls see the halo too
    c->async = saved_async;
    if (rc) { cudaStreamSynchronize(c->stream); return rc; }
    if (!c->async) OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OF3D_OK;
}

OF3D_API int of3d_flow3d_slab(of3d_ctx* c, const void* const* frames_ext, int in_dtype, int64_t nz_ext, int64_t ny, int64_t nx,
                              int64_t own_lo, int64_t own_n, int64_t chunk_planes, const of3d_taps* t, int precision, unsigned flags,
                              void* vx, void* vy, void* vz, void* rel) {
    if (!frames_ext) { set_error("null image or output pointer"); return OF3D_ERR_ARG; }
    return slab_run(c, frames_ext, nullptr, nullptr, in_dtype, nz_ext, ny, nx, own_lo, own_n, chunk_planes, t, precision, flags, vx, vy, vz, rel);
}

OF3D_API int of3d_flow3d_slab_dt(of3d_ctx* c, const void* ic_ext, const void* dt0_ext, int64_t nz_ext, int64_t ny, int64_t nx,
                                 int64_t own_lo, int64_t own_n, int64_t chunk_planes, const of3d_taps* t, int precision, unsigned flags,
                                 void* vx, void* vy, void* vz, void* rel) {
    if (!ic_ext || !dt0_ext) { set_error("null ic/dt0 pointer"); return OF3D_ERR_ARG; }
    return slab_run(c, nullptr, ic_ext, dt0_ext, OF3D_U16, nz_ext, ny, nx, own_lo, own_n, chunk_planes, t, precision, flags, vx, vy, vz, rel);
}

// ---------------------------------------------------------------------------------------------
// TIFF strip decoders (host code; SURVEY.md 8(f) rank 2: the MATLAB twin writes LZW, src/MATLAB/TIFFwrite.m:27).
namespace {
// TIFF 6.0 LZW: MSB-first codes of 9..12 bits, ClearCode 256, EOI 257, the code width grows one code early.
int64_t lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    struct Entry { uint32_t prev; uint32_t len; uint8_t first, last; };
    static thread_local std::vector<Entry> tab(4096);
    for (int i = 0; i < 256; ++i) tab[i] = {0xffffffffu, 1, (uint8_t)i, (uint8_t)i};
    uint32_t next = 258, nbits = 9, acc = 0, have = 0;
    int64_t prev = -1;
    size_t in = 0, out = 0;
    bool cleared = false;
    while (out < cap) {
        while (have < nbits) {
            if (in >= n) return (int64_t)out;
            acc = (acc << 8) | src[in++];
            have += 8;
        }
        const uint32_t code = (acc >> (have - nbits)) & ((1u << nbits) - 1u);
        have -= nbits;
        if (code == 257) break;
        if (code == 256) { next = 258; nbits = 9; prev = -1; cleared = true; continue; }
        if (!cleared) return -1;
        uint32_t emit = code;
        if (prev >= 0) {
            if (code > next || (code == next && next >= 4096)) return -1;
            // new entry = prev + first(code); for code == next (KwKwK) that is prev + first(prev)
            if (next < 4096) tab[next] = {(uint32_t)prev, tab[prev].len + 1, tab[prev].first, code == next ? tab[prev].first : tab[code].first};
            if (next < 4096) ++next;
            if (next + 1 >= (1u << nbits) && nbits < 12) ++nbits;
        } else if (code >= 256) return -1;
        // write the string of `emit` back to front
        const uint32_t len = tab[emit].len;
        size_t end = out + len;
        uint32_t e = emit;
        for (size_t p = end; p-- > out;) {
            if (p < cap) dst[p] = tab[e].last;
            e = tab[e].prev;
        }
        out = end < cap ? end : cap;
        prev = code;
    }
    return (int64_t)out;
}

int64_t packbits_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    size_t in = 0, out = 0;
    while (in < n && out < cap) {
        const int h = (int8_t)src[in++];
        if (h >= 0) {
            size_t cnt = (size_t)h + 1;
            if (in + cnt > n) cnt = n - in;
            for (size_t i = 0; i < cnt && out < cap; ++i) dst[out++] = src[in + i];
            in += cnt;
        } else if (h != -128) {
            if (in >= n) break;
            const uint8_t v = src[in++];
            for (int i = 0; i < 1 - h && out < cap; ++i) dst[out++] = v;
        }
    }
    return (int64_t)out;
}
}  // namespace

OF3D_API int64_t of3d_tiff_decode(int kind, const void* src, size_t n, void* dst, size_t cap) {
    if ((!src && n) || (!dst && cap)) return -1;
    if (kind == 0) return lzw_decode((const uint8_t*)src, n, (uint8_t*)dst, cap);
    if (kind == 1) return packbits_decode((const uint8_t*)src, n, (uint8_t*)dst, cap);
    return -1;
}

OF3D_API int of3d_synth_blobs(of3d_ctx* c, void* dev_out_u16, int64_t nt, int64_t nz, int64_t ny, int64_t nx, int64_t t0, int64_t z0,
                     uint64_t seed) {
    if (!c || !dev_out_u16 || nt < 1 || nz < 1 || ny < 1 || nx < 1) { set_error("bad argument"); return OF3D_ERR_ARG; }
    OF3D_CUDA_TRY(cudaSetDevice(c->device));
    launch_synth_blobs((uint16_t*)dev_out_u16, nt, nz, ny, nx, t0, z0, seed, c->sm_count, c->stream);
    c->launches++;
    OF3D_CUDA_TRY(cudaGetLastError());
    OF3D_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return OF3D_OK;
}

}  // extern "C"
