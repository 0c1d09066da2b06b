import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import generate_mesh
from p1afempy_b200.device import DeviceMesh
level = int(sys.argv[1]) if len(sys.argv) > 1 else 11
c, e = generate_mesh(level)
mesh = DeviceMesh(c, e)
def sync(): torch.cuda.synchronize()
for it in range(6):
    sync(); t0 = time.perf_counter()
    mesh.close(); sync(); t1 = time.perf_counter()
    mesh.plan; sync(); t2 = time.perf_counter()
    out = mesh.stiffness_csr(); sync(); t3 = time.perf_counter()
    print('close %.2f ms  plan %.2f ms  assemble %.2f ms' % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), flush=True)
    del out
print(torch.cuda.memory_allocated()/1e9, torch.cuda.memory_reserved()/1e9)
