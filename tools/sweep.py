#!/usr/bin/env python
"""Config C5: stiffness -> CSR and compute_eta_r over
S(9) .. S(13) (1M .. 256M triangles) on one B200, device-resident inputs and
outputs, cold (nothing cached between calls).  x = nodal interpolant of
sin(pi x) sin(pi y), all-Dirichlet boundary, f = 1 at the centroids.

    python tools/sweep.py > profiles/r01_sweep.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import algorithmic_bytes, mesh_counts
from p1afempy_b200 import meshgen
from p1afempy_b200.device import DeviceMesh


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = []
for level in range(9, 14):
    m, n, nnz = mesh_counts(level)
    c, e, b = meshgen.unit_square_device(level)
    mesh = DeviceMesh(c, e, check=False)
    u = torch.sin(np.pi * c[:, 0]) * torch.sin(np.pi * c[:, 1])
    fc = torch.ones(m, dtype=torch.float64, device=c.device)
    none = torch.zeros((0, 2), dtype=torch.int32, device=c.device)
    reps = 20 if level < 12 else 5
    t_a = timed(lambda: mesh.stiffness_csr(), reps)

    def eta():
        mesh.close()
        return mesh.eta_r(u, b[0], none, None, fc)
    t_e = timed(eta, reps)
    row = {'level': level, 'M': m, 'N': n, 'nnz': nnz,
           'stiffness_to_csr_ms': t_a, 'stiffness_Gelem_s': m / t_a / 1e6,
           'stiffness_algorithmic_GBs': algorithmic_bytes(m, n, nnz) / t_a / 1e6,
           'eta_r_cold_ms': t_e, 'eta_r_Gelem_s': m / t_e / 1e6,
           'eta_r_algorithmic_GBs': 40.0 * m / t_e / 1e6}
    out.append(row)
    print(row, file=sys.stderr, flush=True)
    mesh.close()
    del mesh, c, e, b, u, fc
    torch.cuda.empty_cache()
print(json.dumps({'device': torch.cuda.get_device_name(0), 'rows': out}, indent=1))
