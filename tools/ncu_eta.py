"""One cold pass of the non-matrix operators (load vector, provide_geometric_data,
compute_eta_r) on S(level), for an ncu capture of their kernels:

    ncu --set full --clock-control none --import-source on \
        -k regex:"twin_kernel|eta_|edges_kernel|flags_kernel|gather_nodes|local_load|scatter_kernel" \
        -s <launches of the warm-up pass> -o gpurun_out/eta python tools/ncu_eta.py 12
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import generate_mesh
from bench_ops import meshgen_boundary
from p1afempy_b200 import cubature
from p1afempy_b200.device import DeviceMesh

level = int(sys.argv[1]) if len(sys.argv) > 1 else 12
c, e = generate_mesh(level)
mesh = DeviceMesh(c, e)
w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.CONICAL_GAUSS_4X4)
xq = mesh.quadrature_points(p)
fq = torch.sin(xq[..., 0]) + xq[..., 1]
del xq
u = torch.sin(mesh.coords[:, 0])
fc = torch.ones(mesh.n_elements, dtype=torch.float64, device='cuda')
b = torch.from_numpy(np.vstack(meshgen_boundary(level)).astype(np.int32)).cuda()
none = torch.zeros((0, 2), dtype=torch.int32, device='cuda')
for rep in range(2):            # pass 0 warms the arena, pass 1 is the one to capture
    mesh.close(); mesh.load_vector(fq, w, p)
    mesh.close(); mesh.geometric_data([b])
    mesh.close(); mesh.eta_r(u, b, none, None, fc)
    torch.cuda.synchronize()
print('ok')
