"""Where a small cold assembly spends its time: host enqueue cost per call against the
device time per call, with the arrays reused (graph replay) and with the event
profiler on (plain launches).  python tools/small_probe.py [level]"""
import sys
import time

import torch

sys.path.insert(0, '.')
from p1afempy_b200 import _lib, meshgen            # noqa: E402
from p1afempy_b200.device import DeviceMesh, context  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
c, e, _ = meshgen.unit_square_device(level, device=0)
mesh = DeviceMesh(c, e, device=0)
lib, ctx = _lib.load(), context(0)[0]
for mode in ('replay', 'plain'):
    lib.p1_profile_enable(ctx, 1 if mode == 'plain' else 0)
    cold = None
    for _ in range(10):
        cold = mesh.assemble_cold(_lib.OP_STIFFNESS, out=cold)
    torch.cuda.synchronize()
    n = 300
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        cold = mesh.assemble_cold(_lib.OP_STIFFNESS, out=cold)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print('S(%d) %s: host enqueue %.1f us/call, device %.1f us/call, status %d'
          % (level, mode, (t1 - t0) / n * 1e6, e0.elapsed_time(e1) / n * 1e3,
             int(cold.status[0].item())))
lib.p1_profile_enable(ctx, 0)
