"""CPU tests (gloo, world_size 2 and 3) of the N>1 host logic: timepoint sharding, z-slab planning, the in-place halo
exchange of the raw frames and the owned-range crop.  The compute stage is injected (the oracle, on CPU tensors) -- the
product path uses the CUDA library; what is under test here is that slab + halo + crop reproduces the unsharded
result exactly, including a stack with more frames than the temporal filter touches and a strided input."""
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


def _oracle_flow(ext, own_lo, own_n, sig):
    """The oracle on the extended window, cropped to the owned planes: what of3d_flow3d_slab does on the device."""
    out = orc.lk_flow3d(ext.numpy(), *sig, rel_mode='float64')
    return [torch.from_numpy(np.ascontiguousarray(o[own_lo:own_lo + own_n])) for o in out]


def _worker(rank, world, port, shape, sig, seed, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        img = make_stack(shape, seed=seed, dtype=np.uint16)
        z0, z1 = multigpu.shard_timepoints(shape[1], world)[rank]
        local = torch.from_numpy(img.astype(np.int32))[:, z0:z1]      # a strided (non-contiguous) view of the stack
        out = multigpu.calc_flow3D_zslab(local, *sig, nz_total=shape[1],
                                         flow_fn=lambda ext, lo, n: _oracle_flow(ext, lo, n, sig))
        q.put((rank, [o.numpy() for o in out]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_zslab_halo_exchange_matches_unsharded(world):
    shape, sig, seed = (9, 24, 20, 22), (1, 1, 1.3), 5         # halo = 3 + 4 = 7 planes, 8 planes per rank at world 3; 9 frames, 7 used
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
