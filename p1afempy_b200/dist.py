"""Multi-GPU assembly on one box: one process per GPU (torch.distributed,
NCCL over NVLink/NVSwitch), rows of the matrix split into contiguous ranges of
the ORIGINAL node numbering, so that the per-rank CSR pieces concatenated in
rank order ARE the reference CSR bit for bit.

Design: inputs replicated, owner filter.  Every rank holds (coordinates,
elements) -- 1.3 GB at 64M triangles, nothing next to 180 GB -- and builds the
plan for its own rows only: the count/scatter kernels scan all M elements
(12 B per triangle each) and keep the incidences whose node is owned; records,
row pointer, fill and output scale 1/G.  No partial sums cross GPUs, hence no
data-path collective and results are independent of G; the only collective
is the all-gather of per-rank nnz that turns local row pointers into global
ones (and, on request, the gather of the pieces onto one rank).

The host-side logic below is backend-agnostic (gloo on CPU in the tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition_rows(n_rows: int, world: int):
    """Contiguous, balanced row ranges [(begin, end)] * world."""
    return [(n_rows * r // world, n_rows * (r + 1) // world) for r in range(world)]


def nnz_offsets(local_nnz: int, device=None, group=None):
    """-> (offset of this rank's first entry, total nnz, per-rank nnz)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(local_nnz)], dtype=torch.int64, device=device)
    every = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, mine, group=group)
    every = every.cpu()
    starts = torch.cumsum(every, 0) - every
    return int(starts[rank]), int(every.sum()), [int(x) for x in every]


def gather_csr(indptr, indices, data, n_rows, root=0, group=None):
    """Concatenates the row-range pieces on `root`.

    indptr: local row pointer (rows_owned+1, starting at 0), indices, data:
    torch tensors on the rank's device.  Returns (indptr, indices, data) of the
    whole matrix on root (same device/dtypes), None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = data.device
    offset, total, every = nnz_offsets(data.numel(), device=dev, group=group)
    rows = torch.tensor([indptr.numel() - 1], dtype=torch.int64, device=dev)
    all_rows = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_rows, rows, group=group)
    all_rows = [int(x) for x in all_rows.cpu()]
    assert sum(all_rows) == n_rows, (all_rows, n_rows)
    if rank != root:
        dist.send((indptr[:-1] + offset).contiguous(), dst=root, group=group)
        dist.send(indices.contiguous(), dst=root, group=group)
        dist.send(data.contiguous(), dst=root, group=group)
        return None
    g_indptr = torch.empty(n_rows + 1, dtype=indptr.dtype, device=dev)
    g_indices = torch.empty(total, dtype=indices.dtype, device=dev)
    g_data = torch.empty(total, dtype=data.dtype, device=dev)
    row0, nz0 = 0, 0
    for r in range(world):
        rs, ns = slice(row0, row0 + all_rows[r]), slice(nz0, nz0 + every[r])
        if r == root:
            g_indptr[rs] = indptr[:-1] + offset
            g_indices[ns] = indices
            g_data[ns] = data
        else:
            for buf in (g_indptr[rs], g_indices[ns], g_data[ns]):
                tmp = torch.empty_like(buf)
                dist.recv(tmp, src=r, group=group)
                buf.copy_(tmp)
        row0 += all_rows[r]
        nz0 += every[r]
    g_indptr[n_rows] = total
    return g_indptr, g_indices, g_data


class DistributedMesh:
    """The calling rank's share of a mesh replicated on every GPU."""

    def __init__(self, coordinates, elements, device=None, group=None):
        from .device import DeviceMesh
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        n_nodes = int(coordinates.shape[0])
        self.n_nodes = n_nodes
        self.local = DeviceMesh(coordinates, elements, device=device)
        # ranges balanced by stored entries, not by rows: in NVB numbering the
        # early (coarse-level) nodes have twice the valence of the late ones
        bounds = self.local.balanced_rows(self.world)
        self.rows = (bounds[self.rank], bounds[self.rank + 1])
        self.local.set_rows(self.rows)

    def stiffness_csr(self):
        return self.local.stiffness_csr()

    def mass_csr(self):
        return self.local.mass_csr()

    def offsets(self):
        return nnz_offsets(self.local.nnz, device=self.local.coords.device,
                           group=self.group)

    def gather(self, indptr, indices, data, root=0):
        return gather_csr(indptr, indices, data, self.n_nodes, root=root,
                          group=self.group)

    def close(self):
        self.local.close()
