"""GPU tests of the callers on either side of the hot path: the streaming process_flow driver (TIFF in, TIFF out),
the two-stage C ABI used by the z-slab mode, and (with >= 2 GPUs) the NCCL halo exchange."""
import os
import socket

import numpy as np
import pytest

from oracle import lk_oracle as orc
from opticalflow3d_dev_b200 import tiffio
from opticalflow3d_dev_b200.synth import make_stack

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-9):
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() <= tol * max(np.abs(b).max(), 1e-300)


def test_process_flow_sequence_3d(tmp_path, capsys):
    from opticalflow3d_dev_b200.calc_flow import process_flow
    img = make_stack((9, 10, 40, 44), seed=31, dtype=np.uint16)
    for t in range(9):
        tiffio.imwrite(tmp_path / ('exp_t%03d_ch0.tif' % t), img[t])
    process_flow(str(tmp_path), 'exp_t.*_ch0', 'SequenceT', 3, 1, 1, 2)
    out = tmp_path / 'OpticalFlow3D' / 'exp_t_ch0'
    assert (out / 'exp_t_ch0_parameters.csv').read_text().splitlines() == ['xyzSig,tiSig,wSig,Nx,Ny,Nz,Nt', '1,1,2,44,40,10,9']
    files = sorted(os.listdir(out))
    assert [f for f in files if f.endswith('.tiff')] == sorted(
        'exp_t_ch0_%s_t%04d.tiff' % (n, t) for n in ('vx', 'vy', 'vz', 'rel') for t in (3, 4, 5))
    for c in (3, 4, 5):
        ref = orc.lk_flow3d(img[c - 3:c + 4], 1, 1, 2, rel_mode='float64')
        for k, n in enumerate(('vx', 'vy', 'vz')):
            a = tiffio.imread(out / ('exp_t_ch0_%s_t%04d.tiff' % (n, c)))
            assert a.dtype == np.float64 and a.shape == (10, 40, 44) and _close(a, ref[k])
        rel = tiffio.imread(out / ('exp_t_ch0_rel_t%04d.tiff' % c))
        assert rel.dtype == np.float32 and np.allclose(rel, ref[3], rtol=1e-6, atol=1e-6 * np.abs(ref[3]).max())
    log = capsys.readouterr().out
    assert 'No data will be saved for frame 0 to avoid edge effects' in log and 'Processing frame 3...' in log
    assert 'No data will be saved for frame 8 to avoid edge effects' in log and 'Frame 5 saved.  Duration:' in log


def test_process_flow_frame_step_with_lzw_inputs(tmp_path, capsys):
    """frame_step=2 (plot_FigureS2_dt.m:54) on a SequenceT time-lapse whose files are LZW-compressed multi-page TIFFs
    written by libtiff (the MATLAB twin's own output format, TIFFwrite.m:27)"""
    from PIL import Image
    from opticalflow3d_dev_b200.calc_flow import process_flow
    img = make_stack((15, 4, 24, 40), seed=35, dtype=np.uint16)
    for t in range(15):
        pages = [Image.fromarray(img[t, z]) for z in range(4)]
        pages[0].save(tmp_path / ('lz_t%03d.tif' % t), format='TIFF', save_all=True, append_images=pages[1:], compression='tiff_lzw')
    process_flow(str(tmp_path), 'lz_t.*', 'SequenceT', 3, 1, 1, 2, frame_step=2)
    out = tmp_path / 'OpticalFlow3D' / 'lz_t'
    tiffs = sorted(f for f in os.listdir(out) if f.endswith('.tiff'))
    assert tiffs == sorted('lz_t_%s_t%04d.tiff' % (n, t) for n in ('vx', 'vy', 'vz', 'rel') for t in (6, 7, 8))
    for c in (6, 7, 8):
        ref = orc.lk_flow3d(img[c - 6:c + 7:2], 1, 1, 2, rel_mode='float64')
        for k, n in enumerate(('vx', 'vy', 'vz')):
            assert _close(tiffio.imread(out / ('lz_t_%s_t%04d.tiff' % (n, c))), ref[k])
    log = capsys.readouterr().out
    assert 'No data will be saved for frame 5 to avoid edge effects' in log and 'No data will be saved for frame 9 to avoid edge effects' in log
    assert 'Processing frame 6...' in log and 'Frame 8 saved.' in log


def test_process_flow_onetif_2d(tmp_path):
    from opticalflow3d_dev_b200.calc_flow import process_flow
    img = make_stack((8, 48, 52), seed=32, dtype=np.uint16)
    tiffio.imwrite_imagej(tmp_path / 'movie.tif', img)
    process_flow(tmp_path, 'movie', 'OneTif', 2, 1.5, 1, 2, verbose=False)
    out = tmp_path / 'OpticalFlow2D' / 'movie'
    tiffs = sorted(f for f in os.listdir(out) if f.endswith('.tiff'))
    assert tiffs == sorted('movie_%s_t%04d.tiff' % (n, t) for n in ('vx', 'vy', 'rel') for t in (3, 4))
    for c in (3, 4):
        ref = orc.lk_flow2d(img[c - 3:c + 4], 1.5, 1, 2)
        for k, n in enumerate(('vx', 'vy', 'rel')):
            a = tiffio.imread(out / ('movie_%s_t%04d.tiff' % (n, c)))
            assert a.dtype == np.float64 and _close(a, ref[k])


def test_two_stage_abi_equals_single_call():
    import torch
    from opticalflow3d_dev_b200 import multigpu
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    img = make_stack((7, 12, 40, 36), seed=33, dtype=np.uint16)
    ref = calc_flow3D(img, 3, 1, 4, rel_dtype='float64')
    fr = torch.from_numpy(img).cuda()                           # uint16: the temporal kernels use the fused march's arithmetic
    ic, dt0 = multigpu._cuda_temporal(fr, (3, 1, 4), 'fp64', 0)
    assert np.array_equal(ic.cpu().numpy(), img[3].astype(np.float64))
    outs = multigpu._cuda_spatial(ic, dt0, (3, 1, 4), 'fp64', 0)
    assert all(np.array_equal(o.cpu().numpy(), r) for o, r in zip(outs, ref))
    # wider integer frames take the plain tap-by-tap temporal sum: equal to rounding
    ic3, dt3 = multigpu._cuda_temporal(torch.from_numpy(img.astype(np.int32)).cuda(), (3, 1, 4), 'fp64', 0)
    assert torch.equal(ic3, ic) and float((dt3 - dt0).abs().max()) <= 1e-12 * float(dt0.abs().max())
    # a sliced (non-contiguous) z range of a longer stack gives the same planes; short / even stacks exit like the reference
    pad = np.full((9, 17, 40, 36), 7, np.uint16)
    pad[1:8, 2:14] = img
    ic2, dt2 = multigpu._cuda_temporal(torch.from_numpy(pad).cuda()[:, 2:14], (3, 1, 4), 'fp64', 0)
    assert torch.equal(ic2, ic) and torch.equal(dt2, dt0)
    with pytest.raises(SystemExit):
        multigpu._cuda_temporal(fr[:5], (3, 1, 4), 'fp64', 0)
    with pytest.raises(SystemExit):
        multigpu._cuda_temporal(torch.cat([fr, fr[:1]]), (3, 1, 4), 'fp64', 0)


def _nccl_worker(rank, world, port, shape, sig, seed, chunk, mode, q):
    import torch
    import torch.distributed as dist
    from opticalflow3d_dev_b200 import multigpu
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        img = make_stack(shape, seed=seed, dtype=np.uint16)
        z0, z1 = multigpu.shard_timepoints(shape[1], world)[rank]
        local = torch.from_numpy(img).cuda()[:, z0:z1]                 # strided view: the engine copies it into place
        out = multigpu.calc_flow3D_zslab(local, *sig, nz_total=shape[1], chunk_planes=chunk, exchange=mode)
        q.put((rank, [o.cpu().numpy() for o in out]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('chunk,mode', [(None, 'dt'), (8, 'dt'), (None, 'dt_wide'), (None, 'raw'), (8, 'raw')])
def test_zslab_nccl_two_gpus(chunk, mode):
    """z-slab sharding over two B200s: in-library NCCL halo exchange -- of (ic, dt0) after a boundary-first temporal stage,
    or of the raw frames -- + of3d_flow3d_slab(_dt), bit-identical to the single-GPU result (whole slab, and in chunks
    of 8 planes with the interior chunks ahead of the exchange)."""
    import torch
    import torch.multiprocessing as mp
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    shape, sig, seed = (7, 80, 48, 64), (1, 1, 4), 9           # halo 3 + 12 = 15 planes, 40 planes per rank
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, shape, sig, seed, chunk, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img = make_stack(shape, seed=seed, dtype=np.uint16)
    ref = calc_flow3D(img, *sig, rel_dtype='float64')
    for k in range(4):
        assert np.array_equal(np.concatenate([got[r][k] for r in range(2)], axis=0), ref[k])


@pytest.mark.parametrize('shape,sig,precision,dtype', [
    ((7, 60, 40, 64), (1, 1, 4), 'fp64', np.uint16),       # fused temporal stage (uint16, 16-byte rows), TMA or cp.async window march
    ((7, 60, 33, 50), (1, 1, 4), 'fp64', np.uint16),       # odd rows: separate temporal stage
    ((13, 50, 24, 64), (1, 2, 2), 'fp32', np.uint8),       # 13 frames: run-time temporal loop, uint8
    ((7, 44, 24, 40), (1, 1, 2), 'fp64', np.float32),      # float frames
])
def test_flow3d_slab_ranges_and_chunks_are_bit_identical(shape, sig, precision, dtype):
    """of3d_flow3d_slab on one GPU: any owned z range of an extended window, whole or in chunks, returns exactly the
    planes of the whole-volume call -- the property that makes z-slab sharding (and the slab-pipelined host call) exact.
    The extended window is cut H = R + Rw planes beyond the owned range, as a rank's buffer is after the exchange."""
    import ctypes as C
    import torch
    from opticalflow3d_dev_b200 import _lib, multigpu
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    from opticalflow3d_dev_b200.taps import flow_taps
    img = make_stack(shape, seed=21, dtype=np.uint16).astype(dtype)
    kt, nz, ny, nx = shape
    ref = calc_flow3D(torch.from_numpy(img).cuda(), *sig, precision=precision, rel_dtype='float64')
    H = multigpu.halo_planes(sig[0], sig[2])
    ctx = _lib.get_context(0)
    taps, keep = _lib.make_taps(flow_taps(*sig))
    odt = torch.float64 if precision == 'fp64' else torch.float32
    for (z0, z1, chunk) in [(0, nz, 0), (0, 17, 0), (17, 41, 0), (41, nz, 5), (H + 3, nz - H - 2, 4), (0, nz, 16)]:
        e0, e1 = max(0, z0 - H), min(nz, z1 + H)
        ext = torch.from_numpy(np.ascontiguousarray(img[:, e0:e1])).cuda()
        fb = ext[0].numel() * ext.element_size()
        ptrs = (C.c_void_p * kt)(*[ext.data_ptr() + k * fb for k in range(kt)])
        outs = [torch.empty((z1 - z0, ny, nx), dtype=odt, device='cuda') for _ in range(4)]
        torch.cuda.synchronize()
        rc = ctx.lib.of3d_flow3d_slab(ctx.handle, ptrs, _lib.DTYPE_CODES[np.dtype(dtype)], e1 - e0, ny, nx, z0 - e0, z1 - z0, chunk,
                                      C.byref(taps), _lib.FP64 if precision == 'fp64' else _lib.FP32, 0, *[o.data_ptr() for o in outs])
        _lib.check(rc, 'of3d_flow3d_slab')
        for o, r in zip(outs, ref):
            assert torch.equal(o, r[z0:z1]), (z0, z1, chunk)
        # the same range from extended (ic, dt0) volumes (of3d_temporal + of3d_flow3d_slab_dt): integer frames use the
        # fused march's temporal arithmetic, so the two-stage path is bit-identical as well (float frames: to rounding)
        ic = torch.empty((e1 - e0, ny, nx), dtype=odt, device='cuda')
        dt0 = torch.empty_like(ic)
        outs2 = [torch.empty_like(o) for o in outs]
        torch.cuda.synchronize()
        _lib.check(ctx.lib.of3d_temporal(ctx.handle, 3, ptrs, _lib.DTYPE_CODES[np.dtype(dtype)], _lib.DEVICE, e1 - e0, ny, nx, C.byref(taps),
                                         _lib.FP64 if precision == 'fp64' else _lib.FP32, 0, ic.data_ptr(), dt0.data_ptr()), 'of3d_temporal')
        rc = ctx.lib.of3d_flow3d_slab_dt(ctx.handle, ic.data_ptr(), dt0.data_ptr(), e1 - e0, ny, nx, z0 - e0, z1 - z0, chunk,
                                         C.byref(taps), _lib.FP64 if precision == 'fp64' else _lib.FP32, 0, *[o.data_ptr() for o in outs2])
        _lib.check(rc, 'of3d_flow3d_slab_dt')
        for o, r in zip(outs2, outs):
            if np.dtype(dtype).kind == 'u':
                assert torch.equal(o, r), (z0, z1, chunk, 'dt')
            else:
                assert float((o - r).abs().max()) <= 1e-9 * max(float(r.abs().max()), 1e-30)
    # bad ranges are rejected
    assert ctx.lib.of3d_flow3d_slab(ctx.handle, ptrs, _lib.U16, 10, ny, nx, 5, 6, 0, C.byref(taps), _lib.FP64, 0,
                                    *[o.data_ptr() for o in outs]) == -1


def test_caller_supplied_stream():
    """of3d_set_stream: the library runs on the caller's stream (torch's current stream), so torch work queued before
    and after the call is ordered with it without any host synchronisation."""
    import torch
    from opticalflow3d_dev_b200 import _lib
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    img = make_stack((7, 12, 40, 64), seed=5, dtype=np.uint16)
    ref = calc_flow3D(torch.from_numpy(img).cuda(), 1, 1, 2, rel_dtype='float64')
    ctx = _lib.get_context(0)
    own = ctx.stream
    s = torch.cuda.Stream()
    try:
        _lib.check(ctx.lib.of3d_set_stream(ctx.handle, s.cuda_stream), 'of3d_set_stream')
        assert ctx.stream == s.cuda_stream
        ctx.set_async(True)
        with torch.cuda.stream(s):
            dev = torch.from_numpy(img).cuda(non_blocking=True).clone()   # produced on s, consumed by the library on s
            got = calc_flow3D(dev, 1, 1, 2, rel_dtype='float64')
            tot = [g.sum() for g in got]                                      # torch work after the call, same stream
        s.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(got, ref))
        assert all(torch.equal(t, r.sum()) for t, r in zip(tot, ref))
    finally:
        ctx.set_async(False)
        _lib.check(ctx.lib.of3d_set_stream(ctx.handle, None), 'of3d_set_stream')
    assert ctx.stream == own


def test_flowstream_matches_per_window_calls():
    """The streaming engine (device frame ring, overlapped copies) returns exactly what calc_flow3D returns on each window."""
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    from opticalflow3d_dev_b200.timelapse import FlowStream
    img = make_stack((12, 9, 34, 40), seed=41, dtype=np.uint16)
    eng = FlowStream(img.shape[1:], np.uint16, (1, 1, 2))
    got = {}
    for t in range(img.shape[0]):
        d = eng.push(img[t])
        if d is not None:
            got[d[0]] = [a.copy() for a in d[1]]
    d = eng.flush()
    got[d[0]] = [a.copy() for a in d[1]]
    eng.close()
    assert sorted(got) == [3, 4, 5, 6, 7, 8]
    for c, arrs in got.items():
        ref = calc_flow3D(img[c - 3:c + 4], 1, 1, 2)              # float32 reliability, like the stream's
        assert arrs[3].dtype == np.float32
        assert all(np.array_equal(a, r) for a, r in zip(arrs, ref)), c


def test_flowstream_frame_step_and_copy():
    """frame_step = d: the window of centre c is c + d*(-R..R) (plot_FigureS2_dt.m:54); copy=True returns arrays the
    caller owns; an exception inside the with-block leaves no asynchronous context behind."""
    from opticalflow3d_dev_b200 import _lib
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    from opticalflow3d_dev_b200.timelapse import FlowStream
    img = make_stack((17, 6, 24, 64), seed=43, dtype=np.uint16)
    got = {}
    with FlowStream(img.shape[1:], np.uint16, (1, 1, 2), frame_step=2, copy=True) as eng:
        for t in range(img.shape[0]):
            d = eng.push(img[t])
            if d is not None:
                got[d[0]] = d[1]
        d = eng.flush()
        got[d[0]] = d[1]
    assert sorted(got) == [6, 7, 8, 9, 10]                        # centres 3*2 .. 16 - 3*2
    for c, arrs in got.items():
        ref = calc_flow3D(img[c - 6:c + 7:2], 1, 1, 2)           # strided window, as the MATLAB script builds it
        assert not _lib.is_pinned(arrs[0])
        assert all(np.array_equal(a, r) for a, r in zip(arrs, ref)), c
    with pytest.raises(ValueError):
        with FlowStream(img.shape[1:], np.uint16, (1, 1, 2)) as eng:
            eng.push(img[0])
            eng.push(img[1][:, :5])                               # wrong shape
    out = calc_flow3D(img[:7], 1, 1, 2)                          # the shared context is still synchronous
    assert np.isfinite(out[0]).all()


def test_pinned_buffers_are_released():
    """pinned_empty blocks are freed when the last NumPy view dies (they used to be kept for the process lifetime)"""
    import gc
    import weakref
    from opticalflow3d_dev_b200 import _lib
    a = _lib.pinned_empty((4, 1024), np.float32)
    a[:] = 3.0
    row = a[1]
    base = a
    while getattr(base, 'base', None) is not None:
        base = base.base
    ref = weakref.ref(base)
    del a, base
    gc.collect()
    assert ref() is not None and row[5] == 3.0        # a view keeps the block
    del row
    gc.collect()
    assert ref() is None


def test_stage_times_cover_every_launch():
    """of3d_set_profile / of3d_stage_times: every launch of one operator call lands in exactly one stage bracket"""
    from opticalflow3d_dev_b200 import _lib, calc_flow3D
    ctx = _lib.get_context(0)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 4000, (7, 40, 96, 128)).astype(np.uint16)
    calc_flow3D(img, 3, 1, 4)                                   # warm
    ctx.stage_times()
    ctx.set_profile(True)
    try:
        l0 = ctx.launch_count()
        calc_flow3D(img, 3, 1, 4)
        st = ctx.stage_times()
        assert sum(n for _, n in st.values()) == ctx.launch_count() - l0
        # uint16 frames with 16-byte rows: the temporal derivative is fused into the z march (no 'temporal' stage)
        assert set(st) == {'gradient_xy', 'gradient_z', 'products_window_z', 'window_xy_solve'}
        assert all(ms > 0 for ms, _ in st.values())
        assert ctx.stage_times() == {}                           # cleared
        calc_flow3D(img.astype(np.float32), 3, 1, 4)             # float frames: separate temporal stage
        assert set(ctx.stage_times()) == {'temporal', 'gradient_xy', 'gradient_z', 'products_window_z', 'window_xy_solve'}
        calc_flow3D(img, 3, 1, 4, generic=True, rel_dtype='float64')
        assert set(ctx.stage_times()) == {'temporal', 'generic'}
    finally:
        ctx.set_profile(False)


def test_plain_numpy_call_uses_pooled_pinned_memory():
    """calc_flow3D(pageable ndarray) stages its window through pinned memory and returns arrays in pooled pinned blocks;
    a dead result's block is handed out again; strided / byte-swapped / unsupported-dtype inputs convert on the way."""
    import gc
    from opticalflow3d_dev_b200 import _lib, calc_flow3D
    rng = np.random.default_rng(11)
    img = rng.integers(0, 3000, (9, 6, 40, 64)).astype(np.uint16)
    ref = calc_flow3D(img, 1, 1, 2)
    assert all(_lib.is_pinned(r) for r in ref) and not _lib.is_pinned(img)
    keep = [r.copy() for r in ref]
    del ref
    gc.collect()
    blocks = len(_lib._RANGES)
    again = calc_flow3D(img, 1, 1, 2)
    assert len(_lib._RANGES) == blocks                         # staging and results came back from the pool
    assert all(np.array_equal(a, b) for a, b in zip(again, keep))
    # same values through a strided view, a byte-swapped copy and an int8-free unsupported dtype (uint64 -> float64)
    wide = np.zeros((9, 6, 40, 128), np.uint16); wide[..., ::2] = img
    assert all(np.array_equal(a, b) for a, b in zip(calc_flow3D(wide[..., ::2], 1, 1, 2), keep))
    assert all(np.array_equal(a, b) for a, b in zip(calc_flow3D(img.astype('>u2'), 1, 1, 2), keep))
    # (float64 frames take the separate temporal stage, whose summation order differs from the fused one's by rounding)
    vmax = max(float(np.abs(k).max()) for k in keep[:3])
    for a, b in zip(calc_flow3D(img.astype(np.uint64), 1, 1, 2), keep):
        assert np.abs(a.astype(np.float64) - b).max() <= 1e-9 * max(vmax, float(np.abs(b).max()))
    # a longer stack: only the centre window is staged
    long = np.concatenate([img[:1]] * 2 + [img] + [img[-1:]] * 2)
    assert all(np.array_equal(a, b) for a, b in zip(calc_flow3D(long, 1, 1, 2), keep))
