"""Pinned host<->device copy bandwidth of the box -- the roof of the end-to-end (`e2e`) numbers.

    python tools/pcie_bw.py                       one GPU: D2H alone, H2D alone, both at once
    python tools/pcie_bw.py --gpus 4 [--bind]     k = 1, 2, 4 GPUs copying CONCURRENTLY (one process per GPU): per-GPU and
                                                  aggregate D2H GB/s; --bind pins every process to the CPUs NVML reports as
                                                  local to its GPU before it allocates its page-locked buffer

Prints one JSON line.  The multi-GPU mode answers whether the flat multi-GPU `e2e` curve of bench.py (every rank streams
28 bytes per voxel back over PCIe) is the box or the engine: if the aggregate of k concurrent plain copies stops growing,
no engine can do better on that box.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def one_gpu():
    import torch
    from opticalflow3d_dev_b200 import _lib
    n = 1 << 30
    dev = torch.device('cuda', 0)
    d1, d2 = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    h1 = torch.from_numpy(_lib.pinned_empty((n,), 'uint8')); h2 = torch.from_numpy(_lib.pinned_empty((n,), 'uint8'))
    h1.zero_(); h2.zero_()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(d2h, h2d, reps=5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if d2h:
                with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
            if h2d:
                with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        return reps * n / dt / 1e9

    run(True, True, 1)
    return {'d2h_gbs': run(True, False), 'h2d_gbs': run(False, True), 'both_each_gbs': run(True, True)}


def _worker(rank, k, bind, barrier, q, seconds):
    import torch
    torch.cuda.set_device(rank)
    cpus = None
    if bind:
        from opticalflow3d_dev_b200 import numa
        cpus = numa.bind_to_device(rank)
    from opticalflow3d_dev_b200 import _lib
    n = 1 << 30
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    h = torch.from_numpy(_lib.pinned_empty((n,), 'uint8'))
    h.zero_()
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter(); reps = 0
    while time.perf_counter() - t0 < seconds:
        h.copy_(d, non_blocking=True); torch.cuda.synchronize(); reps += 1
    dt = time.perf_counter() - t0
    q.put((rank, reps * n / dt / 1e9, len(cpus) if cpus else None))
    barrier.wait()


def multi(gpus, bind, seconds=2.0):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = []
    k = 1
    while k <= gpus:
        barrier, q = ctx.Barrier(k), ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, k, bind, barrier, q, seconds)) for r in range(k)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=300) for _ in range(k))
        for p in procs:
            p.join(timeout=60)
        out.append({'gpus_copying': k, 'd2h_gbs_per_gpu': [round(r[1], 2) for r in res], 'd2h_gbs_aggregate': round(sum(r[1] for r in res), 2),
                    'cpus_bound_per_rank': res[0][2]})
        k *= 2
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--bind', action='store_true')
    a = ap.parse_args()
    if a.gpus <= 1:
        print(json.dumps(one_gpu()))
    else:
        print(json.dumps({'host_cpus': os.cpu_count(), 'bind': a.bind, 'concurrent_d2h': multi(a.gpus, a.bind)}))
