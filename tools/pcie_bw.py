"""Pinned host<->device copy bandwidth of the box (the roof of the end-to-end number): D2H alone, H2D alone, both at once."""
import json, sys, time
import torch
sys.path.insert(0, '.')
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
out = {'d2h_gbs': run(True, False), 'h2d_gbs': run(False, True), 'both_each_gbs': run(True, True),
       'torch_pinned_d2h_gbs': None}
hp = torch.empty(n, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): hp.copy_(d1, non_blocking=True)
torch.cuda.synchronize(); out['torch_pinned_d2h_gbs'] = 5 * n / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
