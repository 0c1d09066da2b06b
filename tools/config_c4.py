#!/usr/bin/env python
"""Config C4 (BASELINE.json configs[3]): compute_eta_r on
a >= 64M-triangle ADAPTIVELY refined L-shaped domain, everything device-resident.

Mesh: the reference's L-shape fixture (6 triangles), 10 uniform NVB steps
(6.3M triangles), then adaptive steps -- x = nodal interpolant of
r^(2/3) sin(2 theta/3) about the re-entrant corner (1,1), eta = compute_eta_r(x,
f = 0, all-Dirichlet), Doerfler marking theta = 0.5 (stable descending sort) --
until M >= 67 108 864.  Then compute_eta_r with f = 1 is timed: cold (plan +
twins + estimator, what the reference's stateless function does) and with plan
and twins reused.

    python tools/config_c4.py > profiles/r01_config_c4.json
"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from p1afempy_b200 import meshgen
from p1afempy_b200.device import DeviceMesh

TARGET = int(sys.argv[1]) if len(sys.argv) > 1 else 67_108_864
dev = torch.device('cuda', 0)
c, e, b = meshgen.l_shape()
c, e, b = meshgen.refine_uniform(c, e, b, times=5)
ct = torch.from_numpy(c).to(dev)
et = torch.from_numpy(e.astype(np.int32)).to(dev)
bt = [torch.from_numpy(x.astype(np.int32)).to(dev) for x in b]
for _ in range(5):
    ct, et, bt = meshgen.refine_uniform_device(ct, et, bt)
none = torch.zeros((0, 2), dtype=torch.int32, device=dev)


def singular(xy):
    x, y = xy[:, 0] - 1., xy[:, 1] - 1.
    theta = torch.remainder(torch.atan2(y, x) - math.pi / 2., 2. * math.pi)
    return torch.hypot(x, y) ** (2. / 3.) * torch.sin(2. * theta / 3.)


history = []
t0 = time.perf_counter()
while et.shape[0] < TARGET:
    with DeviceMesh(ct, et, check=False) as mesh:
        eta = mesh.eta_r(singular(ct), torch.cat(bt), none, None,
                         torch.zeros(et.shape[0], dtype=torch.float64, device=dev))
    order = torch.sort(eta, descending=True, stable=True).indices
    csum = torch.cumsum(eta[order], 0)
    k = int(torch.searchsorted(csum, 0.5 * csum[-1])) + 1
    history.append({'M': int(et.shape[0]), 'marked': k, 'eta_sum': float(csum[-1])})
    ct, et, bt = meshgen.refine_nvb_device(ct, et, order[:k].contiguous(), bt)
    del eta, order, csum
torch.cuda.synchronize()
build_s = time.perf_counter() - t0
m, n = int(et.shape[0]), int(ct.shape[0])
area = 0.5 * ((ct[et[:, 1].long(), 0] - ct[et[:, 0].long(), 0]) *
              (ct[et[:, 2].long(), 1] - ct[et[:, 0].long(), 1]) -
              (ct[et[:, 1].long(), 1] - ct[et[:, 0].long(), 1]) *
              (ct[et[:, 2].long(), 0] - ct[et[:, 0].long(), 0]))
x = singular(ct)
ones = torch.ones(m, dtype=torch.float64, device=dev)
dirichlet = torch.cat(bt)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    z.record()
    torch.cuda.synchronize()
    return a.elapsed_time(z) / reps


mesh = DeviceMesh(ct, et, check=False)


def cold():
    mesh.close()
    return mesh.eta_r(x, dirichlet, none, None, ones)


def kernels(fn, reps=3):
    """per-kernel CUDA-event times of one call (library profiler)"""
    import ctypes as C
    from bench import parse_profile
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import context
    ctx, _ = context(0)
    lib = _lib.load()
    fn()
    torch.cuda.synchronize()
    lib.p1_profile_reset(ctx)
    lib.p1_profile_enable(ctx, 1)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 14)
    lib.p1_profile_report(ctx, buf, len(buf))
    lib.p1_profile_enable(ctx, 0)
    return {k: round(v[0] / reps, 3) for k, v in parse_profile(buf.value.decode()).items()}


t_cold = timed(cold)
t_warm = timed(lambda: mesh.eta_r(x, dirichlet, none, None, ones))
eta = mesh.eta_r(x, dirichlet, none, None, ones)
t_asm = timed(lambda: mesh.stiffness_csr())
k_cold = kernels(cold)
k_asm = kernels(lambda: mesh.stiffness_csr())
deg = torch.bincount(et.reshape(-1).long(), minlength=n)
out = {
    'config': 'compute_eta_r on an adaptively refined L-shape (Doerfler 0.5 from 6.3M uniform)',
    'M': m, 'N': n, 'boundary_edges': int(dirichlet.shape[0]),
    'adaptive_steps': len(history), 'mesh_build_s': build_s, 'history': history,
    'area_min': float(area.min()), 'area_max': float(area.max()),
    'levels_spanned': round(math.log(float(area.max() / area.min()), 2), 1),
    'eta_r_cold_ms': t_cold, 'eta_r_cold_Gelem_s': m / t_cold / 1e6,
    'eta_r_reused_ms': t_warm, 'eta_r_reused_Gelem_s': m / t_warm / 1e6,
    'eta_r_reused_algorithmic_GBs': 40.0 * m / t_warm / 1e6,
    'stiffness_to_csr_ms': t_asm, 'stiffness_Gelem_s': m / t_asm / 1e6,
    'eta_sum': float(eta.sum()), 'eta_finite': bool(torch.isfinite(eta).all()),
    'eta_r_cold_kernels_ms': k_cold, 'stiffness_kernels_ms': k_asm,
    'max_elements_per_node': int(deg.max()), 'nodes_with_more_than_8': int((deg > 8).sum()),
}
print(json.dumps(out, indent=1))
