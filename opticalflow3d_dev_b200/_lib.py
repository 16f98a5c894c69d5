"""ctypes binding of libof3d.so (include/of3d.h). There is no CPU fallback: loading fails loudly."""
from __future__ import annotations

import ctypes as C
import weakref
import os
import threading

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, 'libof3d.so')

OK = 0
U8, U16, I16, F32, F64, I32, U32 = range(7)
FP64, FP32 = 0, 1
HOST, DEVICE = 0, 1
FLAG_EXACT, FLAG_GENERIC, FLAG_REL_F32 = 1, 2, 4

DTYPE_CODES = {np.dtype(np.uint8): U8, np.dtype(np.uint16): U16, np.dtype(np.int16): I16,
               np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32,
               np.dtype(np.uint32): U32}


class Taps(C.Structure):
    _fields_ = [(n, t) for k in 'DSGTW' for n, t in ((k, C.POINTER(C.c_double)), ('n' + k, C.c_int32))]


# every symbol include/of3d.h declares: (restype, argtypes)
_vp, _i, _i64, _u, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint, C.c_size_t
SIGNATURES = {
    'of3d_version': (_i, []),
    'of3d_last_error': (C.c_char_p, []),
    'of3d_device_count': (_i, []),
    'of3d_create': (_i, [_i, C.POINTER(_vp)]),
    'of3d_destroy': (_i, [_vp]),
    'of3d_workspace_bytes': (_sz, [_i, _i64, _i64, _i64, _i64, _i, _i, _i, _i]),
    'of3d_reserve': (_i, [_vp, _sz]),
    'of3d_flow3d': (_i, [_vp, _vp, _i, _i, _i64, _i64, _i64, _i64, C.POINTER(Taps), _i, _u, _vp, _vp, _vp, _vp, _i]),
    'of3d_flow2d': (_i, [_vp, _vp, _i, _i, _i64, _i64, _i64, C.POINTER(Taps), _i, _u, _vp, _vp, _vp, _i]),
    'of3d_flow_frames': (_i, [_vp, _i, C.POINTER(_vp), _i, _i, _i64, _i64, _i64, C.POINTER(Taps), _i, _u,
                              _vp, _vp, _vp, _vp, _i]),
    'of3d_flow2d_batch': (_i, [_vp, C.POINTER(_vp), _i, _i64, _i64, _i64, C.POINTER(Taps), _i, _u, _vp, _vp, _vp]),
    'of3d_temporal': (_i, [_vp, _i, C.POINTER(_vp), _i, _i, _i64, _i64, _i64, C.POINTER(Taps), _i, _u, _vp, _vp]),
    'of3d_flow_from_dt': (_i, [_vp, _i, _vp, _vp, _i64, _i64, _i64, C.POINTER(Taps), _i, _u, _vp, _vp, _vp, _vp, _i]),
    'of3d_comm_unique_id': (_i, [_vp]),
    'of3d_comm_init': (_i, [_vp, _vp, _i, _i]),
    'of3d_comm_destroy': (_i, [_vp]),
    'of3d_halo_exchange': (_i, [_vp, C.POINTER(_vp), _i, _sz, _i64, _i64, _i64, _i64, _i64]),
    'of3d_halo_exchange_centre': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i64, _i64, _i64, _i64, _i64, _i64]),
    'of3d_flow3d_slab': (_i, [_vp, C.POINTER(_vp), _i, _i64, _i64, _i64, _i64, _i64, _i64, C.POINTER(Taps), _i, _u,
                              _vp, _vp, _vp, _vp]),
    'of3d_flow3d_slab_dt': (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, C.POINTER(Taps), _i, _u, _vp, _vp, _vp, _vp]),
    'of3d_stream': (_vp, [_vp]),
    'of3d_set_stream': (_i, [_vp, _vp]),
    'of3d_set_async': (_i, [_vp, _i]),
    'of3d_sync': (_i, [_vp]),
    'of3d_launch_count': (_i64, [_vp]),
    'of3d_host_alloc': (_i, [C.POINTER(_vp), _sz]),
    'of3d_host_free': (_i, [_vp]),
    'of3d_window_slab': (_i64, [_i, _i64, _i64, _i64, C.POINTER(Taps)]),
    'of3d_window_upload': (_i, [_vp, _i, _i, _vp, _sz, _sz, _sz]),
    'of3d_window_flow': (_i, [_vp, _i, _i, _i64, _i64, _i64, C.POINTER(Taps), _i, _u, _vp, _vp, _vp, _vp, _i]),
    'of3d_set_profile': (_i, [_vp, _i]),
    'of3d_stage_times': (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_i64)]),
    'of3d_stage_name': (C.c_char_p, [_i]),
    'of3d_order_stats': (_i, [_vp, _vp, _i, _i64, _i64, _i64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
    'of3d_mask_derive': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i64, C.c_double, C.c_double, C.c_double, C.c_double,
                              _vp, _vp, _vp, _vp, _vp, _vp]),
    'of3d_tiff_decode': (_i64, [_i, _vp, _sz, _vp, _sz]),
    'of3d_synth_blobs': (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, C.c_uint64]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libof3d.so (building is __graft_entry__.build()'s job). Raises if it is missing."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError('libof3d.so not found at %s: run `python -m opticalflow3d_dev_b200.build` '
                                   '(this package has no CPU fallback)' % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)      # AttributeError if the symbol is not exported
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def last_error():
    return load().of3d_last_error().decode('utf-8', 'replace')


def check(rc, what):
    if rc != OK:
        raise RuntimeError('%s failed (%d): %s' % (what, rc, last_error()))


def make_taps(tp):
    """dict of float64 arrays -> (Taps struct, keepalive list)."""
    t = Taps()
    keep = []
    for k in 'DSGTW':
        a = np.ascontiguousarray(tp[k], dtype=np.float64)
        keep.append(a)
        setattr(t, k, a.ctypes.data_as(C.POINTER(C.c_double)))
        setattr(t, 'n' + k, a.size)
    return t, keep


class Context:
    """One per (host thread, device): owns the device workspace and stream."""

    def __init__(self, device=0):
        self.lib = load()
        self.device = int(device)
        h = C.c_void_p()
        check(self.lib.of3d_create(self.device, C.byref(h)), 'of3d_create')
        self.handle = h
        self.is_async = False

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.of3d_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self.lib.of3d_stream(self.handle)

    def launch_count(self):
        return int(self.lib.of3d_launch_count(self.handle))

    def sync(self):
        check(self.lib.of3d_sync(self.handle), 'of3d_sync')

    def set_async(self, on):
        check(self.lib.of3d_set_async(self.handle, int(bool(on))), 'of3d_set_async')
        self.is_async = bool(on)

    def set_profile(self, on):
        check(self.lib.of3d_set_profile(self.handle, int(bool(on))), 'of3d_set_profile')

    def stage_times(self):
        """{stage name: (device ms, launches)} since the last call (synchronises the library stream)"""
        n = 6                                                   # OF3D_N_STAGES
        ms, cnt = (C.c_double * n)(), (_i64 * n)()
        check(self.lib.of3d_stage_times(self.handle, ms, cnt), 'of3d_stage_times')
        return {self.lib.of3d_stage_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n) if cnt[i]}


_tls = threading.local()


def get_context(device=None):
    dev = 0 if device is None else int(device)
    cache = getattr(_tls, 'ctx', None)
    if cache is None:
        cache = _tls.ctx = {}
    if dev not in cache:
        cache[dev] = Context(dev)
    return cache[dev]


# ---- page-locked host memory -------------------------------------------------------------------------------------
# Plain NumPy arrays are pageable: copying a 1024x1024x128 result (3.8 GB) into freshly allocated pageable memory takes
# ~700 ms, into pinned memory 67 ms.  Pinned blocks are expensive to create (the pages are locked one by one), so the
# blocks of dead arrays are kept in a size-keyed pool and handed out again.
_POOL_LOCK = threading.RLock()      # re-entrant: a finaliser (_release_block) may run inside a locked region during GC
_POOL = {}                       # rounded size -> [address, ...] of free blocks
_POOL_BYTES = 0
_RANGES = {}                     # address -> size of every live pinned block (pooled or in use)
_POOL_GRAIN = 2 << 20


def _pool_cap():
    return int(float(os.environ.get('OF3D_PINNED_POOL_GB', '24')) * (1 << 30))


def _live_cap():
    """Upper bound on page-locked host memory handed out to callers and the pool together ($OF3D_PINNED_MAX_GB, default
    half of the physical RAM): beyond it pinned_empty(pooled=True) raises and the callers fall back to pageable arrays,
    so a script that keeps many results alive cannot pin the whole machine."""
    env = os.environ.get('OF3D_PINNED_MAX_GB')
    if env is not None:
        return int(float(env) * (1 << 30))
    try:
        return os.sysconf('SC_PAGE_SIZE') * os.sysconf('SC_PHYS_PAGES') // 2
    except (ValueError, OSError):
        return 64 << 30


def _release_block(addr, size, pooled):
    global _POOL_BYTES
    with _POOL_LOCK:
        if pooled and _POOL_BYTES + size <= _pool_cap():
            _POOL.setdefault(size, []).append(addr)
            _POOL_BYTES += size
            return
        _RANGES.pop(addr, None)
    try:
        load().of3d_host_free(C.c_void_p(addr))
    except Exception:
        pass


def pinned_empty(shape, dtype, pooled=False):
    """NumPy array backed by page-locked host memory (cudaHostAlloc) for full-rate PCIe copies.  The block is released
    when the last NumPy view of it dies; pooled=True returns it to a free list instead (capped at $OF3D_PINNED_POOL_GB,
    default 24) so that the next request of the same size costs nothing."""
    global _POOL_BYTES
    lib = load()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(count * dtype.itemsize, 1)
    size = (nbytes + _POOL_GRAIN - 1) // _POOL_GRAIN * _POOL_GRAIN if pooled else nbytes
    addr = None
    if pooled:
        with _POOL_LOCK:
            free = _POOL.get(size)
            if free:
                addr = free.pop()
                _POOL_BYTES -= size
    if addr is None:
        if pooled:
            with _POOL_LOCK:
                live = sum(_RANGES.values())
            if live + size > _live_cap():
                raise RuntimeError('page-locked host memory cap reached (%d bytes live; OF3D_PINNED_MAX_GB)' % live)
        p = C.c_void_p()
        check(lib.of3d_host_alloc(C.byref(p), size), 'of3d_host_alloc')
        addr = p.value
        with _POOL_LOCK:
            _RANGES[addr] = size
    # a Python-level subclass of the ctypes array can carry a finalizer: it runs when the last NumPy view of the block
    # dies (the views keep `buf` alive through their .base chain)
    buf = type('PinnedBuffer', (C.c_char * nbytes,), {}).from_address(addr)
    weakref.finalize(buf, _release_block, addr, size, pooled).atexit = False   # the driver reclaims it at process exit
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def is_pinned(a):
    """True if the NumPy array `a` lies inside a block handed out by pinned_empty."""
    lo = a.ctypes.data
    hi = lo + a.nbytes
    with _POOL_LOCK:
        return any(start <= lo and hi <= start + size for start, size in _RANGES.items())


def pinned_pool_trim():
    """Free every pooled (currently unused) pinned block."""
    global _POOL_BYTES
    with _POOL_LOCK:
        blocks = [(addr, size) for size, lst in _POOL.items() for addr in lst]
        _POOL.clear()
        _POOL_BYTES = 0
        for addr, _ in blocks:
            _RANGES.pop(addr, None)
    for addr, _ in blocks:
        load().of3d_host_free(C.c_void_p(addr))


_COPY_POOL = None


def _copy_pool():
    global _COPY_POOL
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // 2)))
    return _COPY_POOL


def parallel_copy_frames(dst, src, on_frame, parts=4):
    """dst[k] = src[k] for every frame k with several threads, calling on_frame(k) (in order) as soon as frame k is
    complete.  All pieces are queued at once, frame-major, so the workers never idle between frames (a fork-join per
    frame runs at a third of the rate) and the caller can ship frame k while the later ones are still being copied."""
    pool = _copy_pool()
    groups = []
    for k in range(dst.shape[0]):
        n = dst.shape[1] if dst.ndim > 1 else 1
        p = max(1, min(parts, n))
        if dst.ndim < 2 or p == 1:
            groups.append([pool.submit(np.copyto, dst[k:k + 1], src[k:k + 1], 'unsafe')])
        else:
            e = [n * i // p for i in range(p + 1)]
            groups.append([pool.submit(np.copyto, dst[k, a:b], src[k, a:b], 'unsafe') for a, b in zip(e[:-1], e[1:]) if b > a])
    for k, g in enumerate(groups):
        for f in g:
            f.result()
        on_frame(k)


def parallel_copy_pieces(pieces, on_piece):
    """np.copyto(dst, src) for every (dst, src) of `pieces` with several threads, all queued at once; on_piece(i) is
    called in order as soon as piece i is complete (the later ones are still being copied)."""
    pool = _copy_pool()
    futs = [pool.submit(np.copyto, d, s_, 'unsafe') for d, s_ in pieces]
    for i, f in enumerate(futs):
        f.result()
        on_piece(i)


def parallel_copy(dst, src, min_bytes=8 << 20):
    """dst[...] = src with several threads (NumPy releases the GIL inside copyto): one thread moves ~10 GB/s, which is
    less than PCIe.  Both arrays are split along their first axis."""
    if src.nbytes < min_bytes or dst.shape[0] < 2:
        np.copyto(dst, src, casting='unsafe')
        return
    pool = _copy_pool()
    n = dst.shape[0]
    parts = min(n, pool._max_workers)
    edges = [n * i // parts for i in range(parts + 1)]
    futs = [pool.submit(np.copyto, dst[a:b], src[a:b], 'unsafe') for a, b in zip(edges[:-1], edges[1:]) if b > a]
    for f in futs:
        f.result()
