"""How fast a large PAGEABLE numpy array reaches the GPU through device.upload's pinned
staging (threads x chunk size), against torch's own pageable copy and a pinned source.
python tools/upload_probe.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from p1afempy_b200 import device  # noqa: E402

n = 200 * 1000 * 1000                      # 1.6 GB of int64 (the elements of the 64M mesh)
a = np.arange(n, dtype=np.int64)
torch.zeros(1, device='cuda')


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
        del out
    return best


print('torch pageable .to(cuda): %.1f GB/s' % (a.nbytes / timed(lambda: torch.from_numpy(a).to('cuda')) / 1e9))
pinned = torch.empty(n, dtype=torch.int64, pin_memory=True)
pinned.numpy()[:] = a
print('pinned source:            %.1f GB/s' % (a.nbytes / timed(lambda: pinned.to('cuda', non_blocking=True)) / 1e9))
for wave in (2, 4, 6, 8, 12):
    for chunk in (16, 32, 64):
        device._STAGE_WAVE, device._STAGE_CHUNK = wave, chunk << 20
        device._stage.clear()
        t = timed(lambda: device.upload(a, 0))
        print('upload: %2d threads, %2d MB chunks: %.1f GB/s (%.1f ms)' % (wave, chunk, a.nbytes / t / 1e9, t * 1e3))
