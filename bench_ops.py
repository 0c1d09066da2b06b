#!/usr/bin/env python
"""bench_ops.py -- device-resident timing of every operator of the hot path
 on one B200, next to the oracle port on the host for
a bounded sample.  Not the driver's bench (that is bench.py); this fills the
per-operator table in DESIGN.md.

    python bench_ops.py [--level 11] [--cpu-level 9] > profiles/ops.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))


def gpu_time(fn, reps=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_time(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--level', type=int, default=11)
    ap.add_argument('--cpu-level', type=int, default=9)
    args = ap.parse_args()
    import torch
    from p1afempy_b200 import _lib, cubature, meshgen, options
    from p1afempy_b200.device import DeviceMesh
    from oracle import p1_oracle as orc
    import helpers as H

    c, e, _ = meshgen.unit_square(args.level)
    cc, ce, _ = meshgen.unit_square(args.cpu_level)
    m, mc = e.shape[0], ce.shape[0]
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.CONICAL_GAUSS_4X4)
    nq = w.shape[0]
    mesh = DeviceMesh(c, e)
    dev = mesh.coords.device
    out = {'level': args.level, 'M': int(m), 'N': int(c.shape[0]), 'n_qp': int(nq),
           'cpu_level': args.cpu_level, 'M_cpu': int(mc), 'ops': {}}

    def record(name, gpu_ms, cpu_ms, bytes_per_tri):
        out['ops'][name] = {
            'gpu_ms': gpu_ms, 'gpu_Melem_s': m / gpu_ms / 1e3,
            'algorithmic_GBs': bytes_per_tri * m / gpu_ms / 1e6,
            'cpu_ms_sample': cpu_ms, 'cpu_Melem_s': mc / cpu_ms / 1e3 if cpu_ms else None,
            'algorithmic_bytes_per_triangle': bytes_per_tri}

    # device-side analogues of the benchmark callbacks (values at the points)
    xq = mesh.quadrature_points(p)                       # (nq, M, 2)
    x, y = xq[..., 0], xq[..., 1]
    q11 = -(1. + x ** 2); q12 = -0.25 * x * y; q21 = -0.5 * x * y
    q22 = -(2. + torch.sin(np.pi * y))
    fq = torch.sin(2. * np.pi * x) * torch.cos(3. * np.pi * y) + 1.
    u = torch.sin(np.pi * mesh.coords[:, 0]) * torch.sin(np.pi * mesh.coords[:, 1])
    phiq = 1. + mesh.nodal_at_quadrature(u, p) ** 2
    cent = mesh.centroids()
    fc = torch.sin(2. * np.pi * cent[:, 0]) * torch.cos(3. * np.pi * cent[:, 1]) + 1.
    del xq, x, y

    record('stiffness->CSR (cold, no host sync: p1_assemble_cold)',
           gpu_time(lambda: mesh.assemble_cold(_lib.OP_STIFFNESS)),
           cpu_time(lambda: orc.stiffness_coo(cc, ce).tocsr()), 64.0)
    record('stiffness->CSR (cold, plan route)', gpu_time(lambda: mesh.stiffness_csr()), None, 64.0)
    options.plan_flags = _lib.PLAN_TILED
    record('stiffness->CSR (cold, tiled plan: P1_PLAN_TILED)',
           gpu_time(lambda: mesh.stiffness_csr()), None, 64.0)
    options.plan_flags = 0
    reused = DeviceMesh(c, e, reuse=True)
    reused.stiffness_csr()
    record('stiffness->CSR values (DeviceMesh(reuse=True), pattern=False: new values on an '
           'unchanged mesh; cold path without the index arrays)',
           gpu_time(lambda: reused.stiffness_csr(pattern=False)), None,
           8.0 * 3.5 + 12.0 + 8.0)
    reused._cold_refill = False
    record('stiffness->CSR values (the same through p1_assemble_csr on the kept plan: records '
           'refreshed in place, then filled)',
           gpu_time(lambda: reused.stiffness_csr(pattern=False)), None,
           8.0 * 3.5 + 12.0 + 8.0)
    reused.close()
    del reused
    record('mass->CSR (cold, no host sync)', gpu_time(lambda: mesh.assemble_cold(_lib.OP_MASS)),
           cpu_time(lambda: orc.mass_coo(cc, ce).tocsr()), 64.0)
    record('general stiffness->CSR (coefficients array-fed)',
           gpu_time(lambda: mesh.general_stiffness_csr(q11, q12, q21, q22, w)),
           cpu_time(lambda: orc.general_stiffness_coo(cc, ce, H.a11, H.a12, H.a21, H.a22,
                                                      w, p).tocsr(), 1), 64.0 + 32 * nq)
    record('weighted mass->CSR (phi(u_q) array-fed)',
           gpu_time(lambda: mesh.weighted_mass_csr(phiq, w, p)),
           cpu_time(lambda: orc.weighted_mass_coo(cc, ce, H.u_nodal(cc), H.phi, w,
                                                  p).tocsr(), 1), 68.0 + 8 * nq)
    # load vectors: cold = plan (with element ids) + local vectors + gather
    def load_cold():
        mesh.close()
        return mesh.load_vector(fq, w, p)
    record('cubature load vector (f(x_q) array-fed, cold plan)', gpu_time(load_cold),
           cpu_time(lambda: orc.load_vector_cubature(cc, ce, H.f_load, w, p), 1),
           24.0 + 8 * nq)
    record('midpoint load vector (cold)', gpu_time(lambda: mesh.load_vector_midpoint(fc)),
           cpu_time(lambda: orc.load_vector_midpoint(cc, ce, H.f_load)), 32.0)
    kept = DeviceMesh(c, e, reuse=True)          # keeps its (element-keeping) plan
    record('cubature load vector (plan reused)',
           gpu_time(lambda: kept.load_vector(fq, w, p)), None, 24.0 + 8 * nq)
    record('midpoint load vector (plan reused)',
           gpu_time(lambda: kept.load_vector_midpoint(fc)), None, 32.0)
    kept.close()
    # edges + estimator
    b_host = [np.vstack(meshgen_boundary(args.level))]
    bc = [np.vstack(meshgen_boundary(args.cpu_level))]
    b_dev = torch.from_numpy(b_host[0].astype(np.int32)).to(dev)
    none = torch.zeros((0, 2), dtype=torch.int32, device=dev)

    def geo_cold():
        mesh.close()
        return mesh.geometric_data([b_dev])
    record('provide_geometric_data (cold plan)', gpu_time(geo_cold),
           cpu_time(lambda: orc.geometric_data(ce, bc), 1), 36.0)

    def eta_cold():
        mesh.close()
        return mesh.eta_r(u, b_dev, none, None, fc)
    xh = H.u_nodal(cc)
    record('compute_eta_r (cold plan, f at centroids array-fed)', gpu_time(eta_cold),
           cpu_time(lambda: orc.eta_r(xh, cc, ce, bc[0], np.zeros((0, 2), dtype=int),
                                      H.f_load, H.g_neu), 1), 40.0)
    record('compute_eta_r (plan + twins reused)',
           gpu_time(lambda: mesh.eta_r(u, b_dev, none, None, fc)), None, 40.0)
    print(json.dumps(out, indent=1))


def meshgen_boundary(level):
    """Boundary edges of S(level) in the reference's refined order."""
    from p1afempy_b200 import meshgen
    c, e, b = meshgen.criss_square()
    for _ in range(level):
        c, e, b = meshgen.refine_nvb(c, e, np.arange(e.shape[0]), b)
    return b


if __name__ == '__main__':
    main()
