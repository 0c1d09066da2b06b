import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import generate_mesh, parse_profile
from p1afempy_b200 import _lib
from p1afempy_b200.device import DeviceMesh, context
level = int(sys.argv[1]); world = int(sys.argv[2])
c, e = generate_mesh(level)
n = c.shape[0]
ctx, dev = context(0)
lib = _lib.load()
for rank in (0, world // 2, world - 1):
    mesh = DeviceMesh(c, e, interleave=(world, rank, 12))
    for _ in range(3):
        out = mesh.stiffness_csr()
    torch.cuda.synchronize()
    lib.p1_profile_reset(ctx); lib.p1_profile_enable(ctx, 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        out = mesh.stiffness_csr()
    ev1.record(); torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 14)
    lib.p1_profile_report(ctx, buf, len(buf)); lib.p1_profile_enable(ctx, 0)
    prof = parse_profile(buf.value.decode())
    print('world', world, 'rank', rank, 'ms/step %.3f' % (ev0.elapsed_time(ev1) / 10),
          {k: round(v[0] / 10, 3) for k, v in prof.items()}, 'nnz', mesh.nnz, flush=True)
    mesh.close(); del out
