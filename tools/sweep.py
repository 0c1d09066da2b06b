#!/usr/bin/env python
"""Config C5: stiffness -> CSR and compute_eta_r over
S(9) .. S(13) (1M .. 256M triangles), device-resident inputs and outputs, cold
(nothing cached between calls), on 1 GPU or under torchrun on N GPUs of one box
(rows dealt out in chunks of 4096 for the matrix, element blocks for the
estimator; max over ranks).  x = nodal interpolant of sin(pi x) sin(pi y),
all-Dirichlet boundary, f = 1 at the centroids.

    python tools/sweep.py > profiles/r02_sweep_1gpu.json
    python -m torch.distributed.run --nproc-per-node 8 ... tools/sweep.py > ...8gpu.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bench import algorithmic_bytes, mesh_counts
from p1afempy_b200 import _lib, meshgen
from p1afempy_b200 import dist as pdist
from p1afempy_b200.device import DeviceMesh

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
levels = [int(x) for x in sys.argv[1:]] or list(range(9, 14))


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


out = []
for level in levels:
    m, n, nnz = mesh_counts(level)
    c, e, b = meshgen.unit_square_device(level, device=local_rank)
    mesh = DeviceMesh(c, e, device=local_rank, check=False,
                      interleave=(world, rank, 12) if world > 1 else None)
    plain = DeviceMesh(c, e, device=local_rank, check=False)
    u = torch.sin(np.pi * c[:, 0]) * torch.sin(np.pi * c[:, 1])
    fc = torch.ones(m, dtype=torch.float64, device=c.device)
    none = torch.zeros((0, 2), dtype=torch.int32, device=c.device)
    reps = 20 if level < 12 else 5
    statuses = []

    def stiffness():
        cold = mesh.assemble_cold(_lib.OP_STIFFNESS)
        statuses.append(cold.status)
        if world > 1:
            return pdist.global_chunk_offsets(cold.indptr, n, world, rank, 12)
        return cold
    t_a = timed(stiffness, reps)
    assert all(int(s[0]) == 0 for s in statuses)
    statuses.clear()
    # the same call into the SAME output arrays (a time-stepping loop): meshes of at most
    # 4M triangles replay as one CUDA graph
    held = [None]

    def stiffness_again():
        held[0] = mesh.assemble_cold(_lib.OP_STIFFNESS, out=held[0])
        statuses.append(held[0].status.clone())
        if world > 1:
            return pdist.global_chunk_offsets(held[0].indptr, n, world, rank, 12)
        return held[0]
    t_r = timed(stiffness_again, reps)
    assert all(int(s[0]) == 0 for s in statuses)
    held[0] = None
    lo, hi = m * rank // world, m * (rank + 1) // world

    def eta():
        if world > 1:
            return plain.eta_r_block(u, b[0], none, None, fc, lo, hi)
        plain.close()
        return plain.eta_r(u, b[0], none, None, fc)
    t_e = timed(eta, reps)
    row = {'level': level, 'M': m, 'N': n, 'nnz': nnz, 'n_gpus': world,
           'stiffness_to_csr_ms': t_a, 'stiffness_Gelem_s': m / t_a / 1e6,
           'stiffness_to_csr_same_arrays_ms': t_r,
           'stiffness_algorithmic_GBs': algorithmic_bytes('stiffness', m, n, nnz, 0) / t_a / 1e6,
           'eta_r_cold_ms': t_e, 'eta_r_Gelem_s': m / t_e / 1e6,
           'eta_r_algorithmic_GBs': 40.0 * m / t_e / 1e6}
    out.append(row)
    if rank == 0:
        print(row, file=sys.stderr, flush=True)
    mesh.close()
    plain.close()
    del mesh, plain, c, e, b, u, fc, statuses
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({'device': torch.cuda.get_device_name(0), 'n_gpus': world, 'rows': out},
                     indent=1))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
