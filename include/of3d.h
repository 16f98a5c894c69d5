/*
 * of3d.h -- C ABI of libof3d.so: dense Gaussian-weighted Lucas-Kanade optical flow
 * (2D+t and 3D+t) on NVIDIA B200 (sm_100a).
 *
 * The reference (ScientistRachel/OpticalFlow3D_dev) has no FFI: its operator
 * interface for this path is two Python functions.  Each entry point below cites
 * the reference interface it replaces (paths relative to the reference root).
 *
 *   of3d_flow3d  <->  calc_flow3D(images, xyzSig, tSig, wSig)   src/Python/calc_flow.py:175-360
 *   of3d_flow2d  <->  calc_flow2D(images, xySig,  tSig, wSig)   src/Python/calc_flow.py:18-173
 *
 * The sigmas never cross this boundary: the host evaluates the reference's tap
 * expressions (calc_flow.py:72-97 / 230-263) in float64 and passes the five tap
 * vectors, so the doubles are bit-identical to the reference's.
 *
 * Conventions: plain pointers and sizes only; volumes are C-contiguous
 * (t, z, y, x) with x fastest; every function returns OF3D_OK (0) or a negative
 * code and stores a message retrievable with of3d_last_error() (thread-local).
 * Calls are synchronous with respect to the host unless stated otherwise.  One
 * context per (host thread, device); a context is not thread-safe.
 */
#ifndef OF3D_H
#define OF3D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OF3D_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define OF3D_API __attribute__((visibility("default")))
#else
#define OF3D_API
#endif

/* status codes */
#define OF3D_OK 0
#define OF3D_ERR_ARG (-1)     /* bad argument (shape, dtype, taps, null pointer) */
#define OF3D_ERR_CUDA (-2)    /* CUDA runtime error */
#define OF3D_ERR_NOMEM (-3)   /* device or pinned-host allocation failed */
#define OF3D_ERR_NODEVICE (-4)/* no usable CUDA device: there is NO CPU fallback */

/* element type of the input images (calc_flow.py:67,225 widens any real dtype) */
#define OF3D_U8 0
#define OF3D_U16 1
#define OF3D_I16 2
#define OF3D_F32 3
#define OF3D_F64 4
#define OF3D_I32 5
#define OF3D_U32 6

/* arithmetic type of the whole pipeline; outputs are written in the same type */
#define OF3D_FP64 0 /* double: matches the reference to ~1e-12 (bit-exact with OF3D_FLAG_EXACT) */
#define OF3D_FP32 1 /* float filters and windows, fp64 solve; outputs float */

/* where a buffer lives */
#define OF3D_HOST 0
#define OF3D_DEVICE 1

/* flags for of3d_flow3d / of3d_flow2d */
#define OF3D_FLAG_EXACT 1u    /* generic kernels, scipy's paired summation order, no FMA contraction:
                                 fp64 flow fields bit-identical to the reference */
#define OF3D_FLAG_GENERIC 2u  /* force the generic (any tap count) kernels, normal rounding */
#define OF3D_FLAG_REL_F32 4u  /* with OF3D_FP64: `rel` is a float32 buffer, the dtype the reference returns in 3D
                                 (calc_flow.py:355-357); computed in float64, rounded once on the device */

/*
 * The five sampled, un-normalised 1-D filters of the reference, as float64, each of odd length:
 *   D  x*G_sig(x)/sig^2   derivative filter          calc_flow.py:233,235 (fderiv*gderiv)
 *   S  G_{sig/4}(y)       narrow orthogonal smoother calc_flow.py:234      (fsmooth)
 *   G  G_sig(x)           smoother applied to dI/dt  calc_flow.py:253      (fx)
 *   T  t*G_tSig(t)/tSig^2 temporal derivative        calc_flow.py:254,256 (ft*gt)
 *   W  G_wSig(w)          Lucas-Kanade window        calc_flow.py:263      (gw)
 */
typedef struct of3d_taps {
    const double* D; int32_t nD;
    const double* S; int32_t nS;
    const double* G; int32_t nG;
    const double* T; int32_t nT;
    const double* W; int32_t nW;
} of3d_taps;

typedef struct of3d_ctx of3d_ctx; /* per-device workspace + stream */

OF3D_API int of3d_version(void);
OF3D_API const char* of3d_last_error(void);
OF3D_API int of3d_device_count(void);

/* Create / destroy a context on CUDA device `device`.  The context owns a growable device
 * workspace and one stream.  Fails with OF3D_ERR_NODEVICE when no GPU is present. */
OF3D_API int of3d_create(int device, of3d_ctx** out);
OF3D_API int of3d_destroy(of3d_ctx* ctx);

/* Bytes of device workspace one call needs (excluding caller-provided device buffers and,
 * for host inputs/outputs, the staging copies which are included).  ndim is 2 or 3. */
OF3D_API size_t of3d_workspace_bytes(int ndim, int64_t nt_taps, int64_t nz, int64_t ny, int64_t nx, int in_dtype,
                            int precision, int in_mem, int out_mem);
/* Pre-size the workspace (optional; calls grow it on demand). */
OF3D_API int of3d_reserve(of3d_ctx* ctx, size_t bytes);

/*
 * calc_flow3D (calc_flow.py:175-360).
 *   images : (nt, nz, ny, nx) of in_dtype, host or device (in_mem).  Not modified.
 *            nt must be odd and >= taps->nT; the centre frame ceil(nt/2)-1 is analysed
 *            (calc_flow.py:216-223).  Only frames within nT/2 of the centre are read.
 *   vx,vy,vz,rel : (nz, ny, nx) of double (OF3D_FP64) or float (OF3D_FP32), host or device
 *            (out_mem), caller-allocated.  rel = smallest eigenvalue of the windowed
 *            structure tensor (calc_flow.py:352-357), computed in fp64.
 */
OF3D_API int of3d_flow3d(of3d_ctx* ctx, const void* images, int in_dtype, int in_mem,
                int64_t nt, int64_t nz, int64_t ny, int64_t nx,
                const of3d_taps* taps, int precision, unsigned flags,
                void* vx, void* vy, void* vz, void* rel, int out_mem);

/* calc_flow2D (calc_flow.py:18-173).  images (nt, ny, nx); vx, vy, rel (ny, nx). */
OF3D_API int of3d_flow2d(of3d_ctx* ctx, const void* images, int in_dtype, int in_mem,
                int64_t nt, int64_t ny, int64_t nx,
                const of3d_taps* taps, int precision, unsigned flags,
                void* vx, void* vy, void* rel, int out_mem);

/*
 * Same operators on a window given as taps->nT separate frame pointers (all device, or all
 * host), frames[k] = time c - nT/2 + k.  This is what a streaming time-lapse driver
 * (process_flow, calc_flow.py:496-625) calls on its ring of resident frames, so that
 * overlapping windows never re-upload or re-pack frames.  nz = 1 with ndim = 2 for 2D.
 */
OF3D_API int of3d_flow_frames(of3d_ctx* ctx, int ndim, const void* const* frames, int in_dtype, int in_mem,
                     int64_t nz, int64_t ny, int64_t nx,
                     const of3d_taps* taps, int precision, unsigned flags,
                     void* vx, void* vy, void* vz, void* rel, int out_mem);

/*
 * calc_flow2D of `n_out` CONSECUTIVE output timepoints of a 2D time-lapse in one call (the loop of calc_flow.py:599-606):
 * frames[i], i < n_out + nT - 1, are consecutive DEVICE frames; output timepoint j is the centre of frames[j .. j+nT-1];
 * vx, vy, rel are DEVICE buffers (n_out, ny, nx).  The temporal stage runs per timepoint, every later stage once over all
 * n_out planes -- a single 2048x2048 frame does not fill the 148 SMs for more than a wave or two per launch.  Results are
 * identical to n_out separate of3d_flow_frames calls.  n_out + nT - 1 <= 129.
 */
OF3D_API int of3d_flow2d_batch(of3d_ctx* ctx, const void* const* frames, int in_dtype, int64_t n_out, int64_t ny, int64_t nx,
                               const of3d_taps* taps, int precision, unsigned flags, void* vx, void* vy, void* rel);

/*
 * The synchronous host-to-host call (calc_flow3D(ndarray) -> ndarrays), pipelined.  of3d_window_upload copies `bytes`
 * bytes at byte `offset` of frame k (of n_frames, frame_bytes each) from page-locked host memory into a device-resident
 * window on a dedicated copy stream and returns at once, so that the caller can prepare (stage, convert) the next piece
 * while this one crosses PCIe.  A window starts with (k = 0, offset = 0); pieces may come in any order after that.
 * of3d_window_flow runs the operator of of3d_flow_frames on the window.  For large 3D volumes returned to the host it
 * works in z slabs of of3d_window_slab(...) planes: a slab is computed (on its extension by the operator's z support,
 * bit-identical to the whole-volume run) as soon as the planes it needs have arrived -- ship the pieces z-chunk by
 * z-chunk, every frame's chunk before the next chunk -- and copied back while the next slab computes.
 * of3d_window_slab returns 0 when the call is not pipelined.  Replaces the np.double(images) conversion + filtering of
 * calc_flow.py:225-357 / 67-168 for host arrays.
 */
OF3D_API int64_t of3d_window_slab(int ndim, int64_t nz, int64_t ny, int64_t nx, const of3d_taps* taps);
OF3D_API int of3d_window_upload(of3d_ctx* ctx, int k, int n_frames, const void* host_src, size_t frame_bytes, size_t offset,
                                size_t bytes);
OF3D_API int of3d_window_flow(of3d_ctx* ctx, int ndim, int in_dtype, int64_t nz, int64_t ny, int64_t nx,
                              const of3d_taps* taps, int precision, unsigned flags,
                              void* vx, void* vy, void* vz, void* rel, int out_mem);

/*
 * The operator in two stages, for volumes sharded by z-slab across GPUs.  The temporal
 * derivative (calc_flow.py:276-278) is local in z, the spatial stages are not: a rank runs stage 1 on the planes
 * it owns, exchanges R + Rw halo planes of (ic, dt0) with its neighbours (NCCL), runs stage 2 on the extended slab
 * and keeps the planes it owns.
 *   of3d_temporal     frames -> ic (centre frame widened to the compute type) and dt0 (temporal derivative),
 *                     both DEVICE buffers (nz, ny, nx) of double (OF3D_FP64) or float (OF3D_FP32).
 *   of3d_flow_from_dt (ic, dt0) device buffers -> vx, vy, vz, rel (calc_flow.py:279-357 / 116-168).
 */
OF3D_API int of3d_temporal(of3d_ctx* ctx, int ndim, const void* const* frames, int in_dtype, int in_mem,
                           int64_t nz, int64_t ny, int64_t nx, const of3d_taps* taps, int precision, unsigned flags,
                           void* ic_dev, void* dt0_dev);
OF3D_API int of3d_flow_from_dt(of3d_ctx* ctx, int ndim, const void* ic_dev, const void* dt0_dev,
                               int64_t nz, int64_t ny, int64_t nx, const of3d_taps* taps, int precision, unsigned flags,
                               void* vx, void* vy, void* vz, void* rel, int out_mem);

/*
 * z-slab sharding across GPUs, one process per GPU (the z-coupled stages are calc_flow.py:279-313).
 * A rank owns the planes [z0, z1) of the volume and keeps every frame of the window in an EXTENDED buffer: `lo` halo
 * planes below, its `own` planes, `hi` halo planes above, where lo, hi <= H = R + Rw (gradient + window radius) and 0 at
 * the ends of the volume (clamp-to-edge applies there, calc_flow.py:279-313 mode='nearest').
 *
 *   of3d_comm_unique_id / of3d_comm_init / of3d_comm_destroy
 *       the NCCL communicator of the exchange.  Rank 0 obtains a 128-byte id and distributes it by any means (the
 *       Python host broadcasts it with torch.distributed); every rank then calls of3d_comm_init.  NCCL is loaded at
 *       run time (libnccl.so.2), so single-GPU users need none.
 *   of3d_halo_exchange
 *       for each of the n_frames extended frames: sends the `send_dn` lowest owned planes to rank - 1 and the `send_up`
 *       highest to rank + 1, receives `lo` planes from rank - 1 into [0, lo) and `hi` planes from rank + 1 into
 *       [lo + own, lo + own + hi) -- raw planes of the input dtype, straight into place, one grouped ncclSend/ncclRecv on
 *       a dedicated stream.  Returns at once; the next of3d_flow3d_slab waits for it on the device, and only before the
 *       first chunk that touches a halo plane.
 *   of3d_flow3d_slab
 *       calc_flow3D of the planes [own_lo, own_lo + own_n) of the extended window (frames_ext[k]: DEVICE pointers,
 *       nz_ext planes each); vx, vy, vz, rel are DEVICE buffers of own_n planes.  Every stage runs only on the planes the
 *       owned range needs (gradients on own +- Rw, window sums and solve on own).  chunk_planes > 0 works through the
 *       owned range in chunks of that many planes (interior chunks first, overlapping the exchange), which bounds the
 *       workspace by the chunk instead of the slab; 0 = one chunk.  Results are bit-identical to the unsharded call.
 *   of3d_flow3d_slab_dt
 *       the same on extended (ic, dt0) DEVICE volumes of the compute type (of3d_temporal's outputs): a rank runs the
 *       temporal stage on its boundary planes, starts the exchange of those two volumes' halo planes -- 16 (8 in fp32)
 *       bytes per voxel instead of 2 nT, a quarter of the raw frames' halo at nT = 19 -- and runs the temporal stage of
 *       its interior planes while the halos travel.  For 8/16-bit integer frames the temporal kernels use the arithmetic
 *       of the fused z march, so this path, too, is bit-identical to the unsharded call.
 */
OF3D_API int of3d_comm_unique_id(void* id128);
OF3D_API int of3d_comm_init(of3d_ctx* ctx, const void* id128, int nranks, int rank);
OF3D_API int of3d_comm_destroy(of3d_ctx* ctx);
OF3D_API int of3d_halo_exchange(of3d_ctx* ctx, void* const* frames_ext, int n_frames, size_t plane_bytes, int64_t lo, int64_t own,
                                int64_t hi, int64_t send_dn, int64_t send_up);
/* The exchange of the (ic, dt0) halos with ic travelling as what it is the widening of: the raw planes of the centre frame
 * (calc_flow.py:225 `images.astype(np.float64)`; 8/16-bit integer frames): 2 + 8 instead of 8 + 8 bytes per voxel in fp64.
 * `centre_own`: the `own` planes of the centre frame (DEVICE, in_dtype); `stage`: DEVICE scratch of lo + hi raw planes;
 * ic_ext / dt0_ext as in of3d_flow3d_slab_dt.  The planes that arrive are widened into the halo planes of ic_ext on the
 * exchange stream -- the value the neighbour's temporal stage stored, bit for bit. */
OF3D_API int of3d_halo_exchange_centre(of3d_ctx* ctx, const void* centre_own, int in_dtype, void* stage, void* ic_ext, void* dt0_ext,
                                       int precision, int64_t plane_elems, int64_t lo, int64_t own, int64_t hi, int64_t send_dn,
                                       int64_t send_up);
OF3D_API int of3d_flow3d_slab(of3d_ctx* ctx, const void* const* frames_ext, int in_dtype, int64_t nz_ext, int64_t ny, int64_t nx,
                              int64_t own_lo, int64_t own_n, int64_t chunk_planes, const of3d_taps* taps, int precision,
                              unsigned flags, void* vx, void* vy, void* vz, void* rel);
OF3D_API int of3d_flow3d_slab_dt(of3d_ctx* ctx, const void* ic_ext, const void* dt0_ext, int64_t nz_ext, int64_t ny, int64_t nx,
                                 int64_t own_lo, int64_t own_n, int64_t chunk_planes, const of3d_taps* taps, int precision,
                                 unsigned flags, void* vx, void* vy, void* vz, void* rel);

/* Stream control: the context's stream as a cudaStream_t (for CUDA-event timing by the caller), a caller-supplied
 * stream (of3d_set_stream: every later kernel and copy of the context runs on `cuda_stream`, a cudaStream_t of the
 * context's device; NULL restores the context's own stream; the SURVEY 8(b) sketch's `void* cuda_stream` argument),
 * asynchronous mode (device in/out only: calls return after enqueueing), and a sync. */
OF3D_API void* of3d_stream(of3d_ctx* ctx);
OF3D_API int of3d_set_stream(of3d_ctx* ctx, void* cuda_stream);
OF3D_API int of3d_set_async(of3d_ctx* ctx, int enable);
OF3D_API int of3d_sync(of3d_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
OF3D_API int64_t of3d_launch_count(of3d_ctx* ctx);

/*
 * Per-stage device time (measurement hook, nothing in the reference): with profiling on, every kernel launch of the
 * flow operators is bracketed by two CUDA events on the context's stream.  of3d_stage_times synchronises, adds the
 * elapsed time (ms) and launch count of every bracket since the last call into ms[OF3D_N_STAGES] /
 * launches[OF3D_N_STAGES] (either may be null) and clears the accumulators.
 */
#define OF3D_STAGE_TEMPORAL 0         /* calc_flow.py:276-278  temporal derivative of the centre frame          */
#define OF3D_STAGE_GRAD_XY 1          /* calc_flow.py:279-312  in-plane passes of dx, dy, dz, dt                */
#define OF3D_STAGE_GRAD_Z 2           /* calc_flow.py:276-312  z passes of dx, dy, dz, dt (+ fused temporal derivative) */
#define OF3D_STAGE_WINDOW_Z 3         /* calc_flow.py:315-331  nine products and the z pass of the window       */
#define OF3D_STAGE_WINDOW_XY_SOLVE 4  /* calc_flow.py:315-357  in-plane window passes, solve, reliability       */
#define OF3D_STAGE_GENERIC 5          /* the same lines on the generic (any tap count / bit-exact) kernels      */
#define OF3D_N_STAGES 6
OF3D_API int of3d_set_profile(of3d_ctx* ctx, int enable);
OF3D_API int of3d_stage_times(of3d_ctx* ctx, double* ms, int64_t* launches);
OF3D_API const char* of3d_stage_name(int stage);

/* Pinned host buffers for callers that want full-rate host<->device copies. */
OF3D_API int of3d_host_alloc(void** ptr, size_t bytes);
OF3D_API int of3d_host_free(void* ptr);

/*
 * Downstream analysis step (reference src/Python/example_analysis_script.ipynb cells 4-6), on device buffers:
 *   of3d_order_stats  the k_lo-th and k_hi-th smallest elements (0-based ranks among the n values) of a float / double
 *                     array by radix select: what np.percentile(rel, relPer) interpolates between (cell 4).  *n_nan
 *                     receives the number of NaNs; when it is non-zero NumPy's answer is NaN and the ranks are not computed.
 *   of3d_mask_derive  relMask = rel > thresh; v = v * relMask; v[v == 0] = nan; v = v * scale / tscale (xyscale for
 *                     vx, vy; zscale for vz); Magnitude = sqrt(vx^2+vy^2+vz^2); theta = arctan2(vy, vx);
 *                     phi = arctan(vz / sqrt(vx^2+vy^2)) (cells 5-6), one fused pass.  vz/oz/phi may be null (2D).
 *                     Outputs have the dtype of the velocities.
 */
OF3D_API int of3d_order_stats(of3d_ctx* ctx, const void* data_dev, int is_f64, int64_t n, int64_t k_lo, int64_t k_hi,
                              double* out_lo, double* out_hi, int64_t* n_nan);
OF3D_API int of3d_mask_derive(of3d_ctx* ctx, const void* vx, const void* vy, const void* vz, const void* rel, int v_f64, int rel_f64,
                              int64_t n, double thresh, double xyscale, double zscale, double tscale,
                              void* ox, void* oy, void* oz, void* mag, void* theta, void* phi);

/*
 * TIFF strip / tile decoders for the time-lapse driver's reader (host code, no GPU; the reference reads through tifffile,
 * calc_flow.py:445-465,509,571, and its MATLAB twin writes LZW, src/MATLAB/TIFFwrite.m:27).  kind 0 = TIFF 6.0 LZW,
 * 1 = PackBits.  Decodes at most `cap` bytes into dst; returns the number of bytes produced, or -1 for a corrupt stream.
 */
OF3D_API int64_t of3d_tiff_decode(int kind, const void* src, size_t n, void* dst, size_t cap);

/*
 * Benchmark utility: fill a device buffer (nt, nz, ny, nx) of uint16 with the synthetic
 * translating / deforming Gaussian-blob model of SURVEY.md 8(d) (one blob per 16^3 cell,
 * counter-based hashing, offset 100, Gaussian noise sigma 5).  t0 is the time index of the
 * first frame so that shards of one time-lapse can be generated independently; z0 likewise.
 */
OF3D_API int of3d_synth_blobs(of3d_ctx* ctx, void* dev_out_u16, int64_t nt, int64_t nz, int64_t ny, int64_t nx,
                     int64_t t0, int64_t z0, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* OF3D_H */
