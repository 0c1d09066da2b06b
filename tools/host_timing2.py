import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import generate_mesh
from p1afempy_b200.device import DeviceMesh
level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
keep = len(sys.argv) > 2 and sys.argv[2] == 'keep'
c, e = generate_mesh(level)
mesh = DeviceMesh(c, e)
def sync(): torch.cuda.synchronize()
out = None
for it in range(8):
    sync(); t0 = time.perf_counter()
    mesh.close(); t1 = time.perf_counter()
    mesh.plan; t2 = time.perf_counter()
    new = mesh.stiffness_csr(); t3 = time.perf_counter()
    if keep: out = new
    del new
    sync(); t4 = time.perf_counter()
    print('close %.2f  plan %.2f  assemble %.2f  tail %.2f ms' % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3), flush=True)
print(torch.cuda.memory_allocated()/1e9, torch.cuda.memory_reserved()/1e9)
