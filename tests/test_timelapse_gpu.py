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
    fr = torch.from_numpy(img.astype(np.int32)).cuda()
    ic, dt0 = multigpu._cuda_temporal(fr, (3, 1, 4), 'fp64', 0)
    assert np.array_equal(ic.cpu().numpy(), img[3].astype(np.float64))
    outs = multigpu._cuda_spatial(ic, dt0, (3, 1, 4), 'fp64', 0)
    assert all(np.array_equal(o.cpu().numpy(), r) for o, r in zip(outs, ref))


def _nccl_worker(rank, world, port, shape, sig, seed, q):
    import torch
    import torch.distributed as dist
    from opticalflow3d_dev_b200 import multigpu
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        img = make_stack(shape, seed=seed, dtype=np.uint16)
        z0, z1 = multigpu.shard_timepoints(shape[1], world)[rank]
        local = torch.from_numpy(img[:, z0:z1].astype(np.int32)).cuda()
        out = multigpu.calc_flow3D_zslab(local, *sig, nz_total=shape[1])
        q.put((rank, [o.cpu().numpy() for o in out]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_zslab_nccl_two_gpus():
    import torch
    import torch.multiprocessing as mp
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    shape, sig, seed = (7, 40, 48, 52), (1, 1, 4), 9           # halo 3 + 12 = 15 planes, 20 planes per rank
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, shape, sig, seed, q)) for r in range(2)]
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
        assert set(st) == {'temporal', 'gradient_xy', 'gradient_z', 'products_window_z', 'window_xy_solve'}
        assert all(ms > 0 for ms, _ in st.values())
        assert ctx.stage_times() == {}                           # cleared
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
    assert all(np.array_equal(a, b) for a, b in zip(calc_flow3D(img.astype(np.uint64), 1, 1, 2), keep))
    # a longer stack: only the centre window is staged
    long = np.concatenate([img[:1]] * 2 + [img] + [img[-1:]] * 2)
    assert all(np.array_equal(a, b) for a, b in zip(calc_flow3D(long, 1, 1, 2), keep))
