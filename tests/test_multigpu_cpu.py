"""CPU tests (gloo, world_size 2 and 3) of the N>1 host logic: timepoint sharding, z-slab planning, halo
exchange and cropping.  The compute stages are injected (the oracle, on CPU tensors) -- the product path uses
the CUDA library; what is under test here is that slab + halo + crop reproduces the unsharded result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lk_oracle as orc
from opticalflow3d_dev_b200 import multigpu
from opticalflow3d_dev_b200.synth import make_stack
from opticalflow3d_dev_b200.timelapse import shard_range


def test_shard_timepoints_cover_and_balance():
    for n in (0, 1, 7, 55):
        for w in (1, 2, 3, 8):
            s = multigpu.shard_timepoints(n, w)
            assert s[0][0] == 0 and s[-1][1] == n and all(a[1] == b[0] for a, b in zip(s, s[1:]))
            sizes = [b - a for a, b in s]
            assert max(sizes) - min(sizes) <= 1
            assert [shard_range(n, r, w) for r in range(w)] == s
    assert multigpu.shard_timepoints(55, 8)[0] == (0, 7)      # cfg4: 7 vs 6.875 timepoints per GPU


def test_plan_slabs():
    h = multigpu.halo_planes(3, 8)
    assert h == 9 + 24
    p = multigpu.plan_slabs(512, 8, h)
    assert p[0]['own'] == (0, 64) and p[0]['lo'] == 0 and p[0]['hi'] == 33 and p[3]['ext'] == (192 - 33, 256 + 33)
    assert p[7]['hi'] == 0
    with pytest.raises(ValueError):
        multigpu.plan_slabs(64, 8, h)                         # 8 planes per rank < 33-plane halo


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _oracle_temporal(fr, sig):
    img = fr.numpy().astype(np.float64)
    tp = orc.make_taps(*sig)
    c = img.shape[0] // 2
    rt = tp['T'].size // 2
    dt0 = orc.correlate1d_nearest(img[c - rt:c + rt + 1], tp['T'], 0)[rt]
    return torch.from_numpy(np.ascontiguousarray(img[c])), torch.from_numpy(np.ascontiguousarray(dt0))


def _oracle_spatial(ic, dt0, sig):
    """Spatial stages of the oracle on (ic, dt0): same code path as lk_flow3d after the temporal stage."""
    tp = orc.make_taps(*sig)
    D, S, G, W = tp['D'], tp['S'], tp['G'], tp['W']
    corr = orc.correlate1d_nearest
    ic, dt0 = ic.numpy(), dt0.numpy()
    ch = lambda a, f: orc._chain(a, [(f[0], 1), (f[1], 2), (f[2], 0)], corr)
    dt, dy, dx, dz = ch(dt0, (G, G, G)), ch(ic, (D, S, S)), ch(ic, (S, D, S)), ch(ic, (S, S, D))
    win = lambda p: ch(p, (W, W, W))
    tx, ty, tz = win(dx * dt), win(dy * dt), win(dz * dt)
    xy, xz, xx, yz, yy, zz = win(dx * dy), win(dx * dz), win(dx * dx), win(dy * dz), win(dy * dy), win(dz * dz)
    det = (xx * yy * zz) + (2 * xy * xz * yz) - (yy * xz ** 2) - (zz * xy ** 2) - (xx * yz ** 2)
    inv = (det + orc.EPS) ** -1
    vx = -inv * ((yy * zz - yz * yz) * tx + (xz * yz - xy * zz) * ty + (xy * yz - xz * yy) * tz)
    vy = -inv * ((yz * xz - xy * zz) * tx + (xx * zz - xz * xz) * ty + (xz * xy - xx * yz) * tz)
    vz = -inv * ((xy * yz - yy * xz) * tx + (xy * xz - xx * yz) * ty + (xx * yy - xy * xy) * tz)
    rel = orc.min_eig_sym3(xx, xy, xz, yy, yz, zz, 'float64')
    return [torch.from_numpy(np.ascontiguousarray(a)) for a in (vx, vy, vz, rel)]


def _worker(rank, world, port, shape, sig, seed, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        img = make_stack(shape, seed=seed, dtype=np.uint16)
        z0, z1 = multigpu.shard_timepoints(shape[1], world)[rank]
        local = torch.from_numpy(img[:, z0:z1].astype(np.int32))
        out = multigpu.calc_flow3D_zslab(local, *sig, nz_total=shape[1],
                                         temporal_fn=lambda fr: _oracle_temporal(fr, sig),
                                         spatial_fn=lambda a, b: _oracle_spatial(a, b, sig))
        q.put((rank, [o.numpy() for o in out]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_zslab_halo_exchange_matches_unsharded(world):
    shape, sig, seed = (7, 24, 20, 22), (1, 1, 1.3), 5         # halo = 3 + 4 = 7 planes, 8 planes per rank at world 3
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, sig, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img = make_stack(shape, seed=seed, dtype=np.uint16)
    ref = orc.lk_flow3d(img, *sig, rel_mode='float64')
    for k in range(4):
        full = np.concatenate([got[r][k] for r in range(world)], axis=0)
        assert np.array_equal(full, ref[k]), k                 # bit-identical: same arithmetic on every owned plane
