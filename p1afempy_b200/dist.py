"""Multi-GPU assembly on one box: one process per GPU (torch.distributed,
NCCL over NVLink/NVSwitch).  The rows of the matrix are dealt out to the ranks
in chunks of 2^chunk_log2 consecutive rows of the ORIGINAL node numbering
(chunk c belongs to rank c % world), so that the ranks' CSR pieces interleaved
chunk by chunk ARE the reference CSR bit for bit, and
every rank holds the same mix of old (high-valence) and new nodes.

Design: inputs replicated, owner filter.  Every rank holds (coordinates,
elements) -- 1.3 GB at 64M triangles, nothing next to 180 GB -- and builds the
plan for its own rows only: the scatter kernel scans all M elements (12 B per
triangle) and keeps, and computes, only the incidences whose node is owned;
records, row pointer, fill and output scale 1/G.  No partial sums cross GPUs,
hence no data-path collective and results are independent of G; the only
collectives are size exchanges and, on request, the gather of the pieces onto
one rank.

The host-side logic below is backend-agnostic (gloo on CPU in the tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

DEFAULT_CHUNK_LOG2 = 12


def partition_rows(n_rows: int, world: int):
    """Contiguous, balanced row ranges [(begin, end)] * world (by row count)."""
    return [(n_rows * r // world, n_rows * (r + 1) // world) for r in range(world)]


def interleaved_rows(n_rows: int, world: int, rank: int, chunk_log2: int, device=None):
    """Global row ids owned by `rank`: chunks rank, rank+world, ... in order."""
    chunks = torch.arange(rank, max(rank, (n_rows + (1 << chunk_log2) - 1) >> chunk_log2), world,
                          dtype=torch.int64, device=device)
    rows = (chunks[:, None] << chunk_log2) + torch.arange(1 << chunk_log2, dtype=torch.int64,
                                                          device=device)
    rows = rows.reshape(-1)
    return rows[rows < n_rows]


def nnz_offsets(local_nnz: int, device=None, group=None):
    """-> (sum of nnz of lower ranks, total nnz, per-rank nnz)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(local_nnz)], dtype=torch.int64, device=device)
    every = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, mine, group=group)
    every = every.cpu()
    starts = torch.cumsum(every, 0) - every
    return int(starts[rank]), int(every.sum()), [int(x) for x in every]


def global_chunk_offsets(indptr, n_rows, world, rank, chunk_log2, group=None):
    """Interleaved ownership: where every chunk of the calling rank starts in the
    GLOBAL CSR -- what places a rank's piece in the global matrix.

    indptr: the rank's local row pointer (owned rows + 1, starting at 0, CUDA or
    CPU tensor).  Every rank contributes the nnz of each of its chunks (one
    all-gather of n_rows / 2^chunk_log2 / world integers, the only collective of a
    multi-GPU assembly); an exclusive scan over the chunks in global order gives
    every chunk its offset.  Returns (offsets of my chunks, int64; total nnz as a
    0-dim tensor).  The global position of local row a is
    ``offsets[a >> chunk_log2] + indptr[a] - indptr[(a >> chunk_log2) << chunk_log2]``
    (global_row_pointer materialises it).  A few kilobytes of tensors, everything on
    the device, nothing synchronises the host."""
    dev = indptr.device
    c = 1 << chunk_log2
    n_chunks = (n_rows + c - 1) >> chunk_log2
    per_rank = (n_chunks + world - 1) // world            # padded: same length on every rank
    n_own = indptr.numel() - 1
    every = torch.empty(world * per_rank, dtype=torch.int64, device=dev)
    if dev.type == 'cuda':       # two launches around the collective (csrc/afem.cu)
        from . import _lib
        from .device import _ptr, _stream, context
        ctx, _ = context(dev.index)
        lib = _lib.load()
        mine = torch.empty(per_rank, dtype=torch.int64, device=dev)
        _lib.check(lib.p1_chunk_nnz(ctx, _ptr(indptr), n_own, chunk_log2, per_rank, _ptr(mine),
                                    _stream(dev.index)))
        dist.all_gather_into_tensor(every, mine, group=group)
        offs = torch.empty(per_rank + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.p1_chunk_offsets(ctx, _ptr(every), world, per_rank, rank, _ptr(offs),
                                        _ptr(offs[per_rank:]), _stream(dev.index)))
        return offs[:per_rank], offs[per_rank]
    starts = torch.arange(0, per_rank * c, c, dtype=torch.int64, device=dev).clamp_(max=n_own)
    ends = (starts + c).clamp_(max=n_own)
    mine = (indptr[ends] - indptr[starts]).to(torch.int64)   # nnz of my chunks (0: padding)
    dist.all_gather_into_tensor(every, mine, group=group)
    in_order = every.reshape(world, per_rank).t().reshape(-1)   # chunk q * world + r
    offs = torch.cumsum(in_order, 0) - in_order
    return offs.reshape(per_rank, world)[:, rank].contiguous(), in_order.sum()


def global_row_pointer(indptr, n_rows, world, rank, chunk_log2, group=None):
    """-> (global start of every owned row, int64; total nnz): global_chunk_offsets
    spread over the rows."""
    offs, total = global_chunk_offsets(indptr, n_rows, world, rank, chunk_log2, group=group)
    n_own = indptr.numel() - 1
    ip = indptr.to(torch.int64)
    q = torch.arange(n_own, dtype=torch.int64, device=indptr.device) >> chunk_log2
    return offs[q] + ip[:-1] - ip[q << chunk_log2], total


def _place(g_indptr, g_indices, g_data, rows, indptr, indices, data):
    """Copies one piece (local CSR over the global rows `rows`) into place."""
    lens = (indptr[1:] - indptr[:-1]).to(torch.int64)
    start = g_indptr[rows].to(torch.int64)
    dest = torch.repeat_interleave(start - indptr[:-1].to(torch.int64), lens) + \
        torch.arange(indices.numel(), dtype=torch.int64, device=indices.device)
    g_indices[dest] = indices
    g_data[dest] = data


def gather_csr(indptr, indices, data, rows, n_rows, root=0, group=None):
    """Assembles the whole matrix on `root` from the ranks' pieces.

    indptr: local row pointer (len(rows)+1, starting at 0); rows: the global
    row id of every local row (any disjoint cover of range(n_rows) over the
    ranks: contiguous ranges or interleaved chunks).  Returns (indptr, indices,
    data) on root (same device/dtypes), None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = data.device
    sizes = torch.tensor([rows.numel(), data.numel()], dtype=torch.int64, device=dev)
    every = torch.zeros(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(every, sizes, group=group)
    every = every.cpu().reshape(world, 2)
    assert int(every[:, 0].sum()) == n_rows, (every, n_rows)
    total = int(every[:, 1].sum())
    if rank != root:
        for t in (rows, indptr, indices, data):
            dist.send(t.contiguous(), dst=root, group=group)
        return None
    pieces = []
    for r in range(world):
        if r == root:
            pieces.append((rows, indptr, indices, data))
            continue
        nr, nz = int(every[r, 0]), int(every[r, 1])
        bufs = (torch.empty(nr, dtype=rows.dtype, device=dev),
                torch.empty(nr + 1, dtype=indptr.dtype, device=dev),
                torch.empty(nz, dtype=indices.dtype, device=dev),
                torch.empty(nz, dtype=data.dtype, device=dev))
        for t in bufs:
            dist.recv(t, src=r, group=group)
        pieces.append(bufs)
    lens = torch.zeros(n_rows, dtype=torch.int64, device=dev)
    for prow, pptr, _, _ in pieces:
        lens[prow] = (pptr[1:] - pptr[:-1]).to(torch.int64)
    g_indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=g_indptr[1:])
    g_indices = torch.empty(total, dtype=indices.dtype, device=dev)
    g_data = torch.empty(total, dtype=data.dtype, device=dev)
    for piece in pieces:
        _place(g_indptr, g_indices, g_data, *piece)
    return g_indptr.to(indptr.dtype), g_indices, g_data


class DistributedMesh:
    """The calling rank's share of a mesh replicated on every GPU."""

    def __init__(self, coordinates, elements, device=None, group=None,
                 chunk_log2=DEFAULT_CHUNK_LOG2):
        from .device import DeviceMesh
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_nodes = int(coordinates.shape[0])
        self.local = DeviceMesh(coordinates, elements, device=device,
                                interleave=(self.world, self.rank, chunk_log2))

        self.chunk_log2 = chunk_log2

    # Every method returns the calling rank's rows (local CSR: indptr starting at 0)
    # or its entries of a nodal vector, in the order of global_rows().  Matrices
    # take the cold path without a host synchronisation (p1_assemble_cold).
    def _cold(self, op, **kw):
        return self.local.assemble_cold(op, **kw).wait()

    def stiffness_csr(self):
        from . import _lib
        return self._cold(_lib.OP_STIFFNESS)

    def mass_csr(self):
        from . import _lib
        return self._cold(_lib.OP_MASS)

    def general_stiffness_csr(self, a11q, a12q, a21q, a22q, weights):
        """a_ij_q: (n_qp, M) values of the coefficients at the quadrature points of
        ALL elements (replicated like the mesh); the rank evaluates the local matrices
        of the elements that touch its rows inside the scatter (solvers.py:278-377)."""
        from . import _lib
        from .device import _host_f64, to_device_f64
        w = _host_f64(weights)
        qs = [to_device_f64(a, self.local.device, (w.shape[0], self.local.n_elements))
              for a in (a11q, a12q, a21q, a22q)]
        args = _lib.QuadArgs(qs[0].data_ptr(), qs[1].data_ptr(), qs[2].data_ptr(),
                             qs[3].data_ptr(), None, w.ctypes.data, None, w.shape[0])
        return self._cold(_lib.OP_GENERAL, quad=args, keepalive=(qs, w))

    def weighted_mass_csr(self, phiq, weights, points):
        """phiq: (n_qp, M) values of phi(u_h) at the quadrature points
        (solvers.py:50-142)."""
        from . import _lib
        from .device import _host_f64, to_device_f64
        w, p = _host_f64(weights), _host_f64(points)
        phiq = to_device_f64(phiq, self.local.device, (w.shape[0], self.local.n_elements))
        args = _lib.QuadArgs(None, None, None, None, phiq.data_ptr(), w.ctypes.data,
                             p.ctypes.data, w.shape[0])
        return self._cold(_lib.OP_WEIGHTED, quad=args, keepalive=(phiq, w, p))

    def load_vector(self, fq, weights, points):
        """Entries of the cubature load vector (solvers.py:540-598) at the owned rows."""
        return self.local.load_vector(fq, weights, points)

    def load_vector_midpoint(self, f_centroid):
        """Entries of the legacy load vector (solvers.py:414-426) at the owned rows."""
        return self.local.load_vector_midpoint(f_centroid)

    # The two methods above repeat the element loop on every rank (each rank needs the
    # local loads of every element that touches its rows: nearly all of them), which is
    # bit-identical to one GPU but does not scale.  The *_reduced forms deal out the
    # ELEMENTS instead -- SURVEY.md section 8(e): "for vectors ncclReduceScatter" --: the
    # rank assembles the vector of its contiguous element block over all N nodes and one
    # NCCL all-reduce adds the ranks' partial vectors.  The sums are associated by block,
    # so the result agrees with the reference to rounding (<= 1e-12 of the sum of the
    # absolute contributions; the contract for vectors), not bit for bit, and depends on
    # the number of ranks in the last bits.
    def _block_mesh(self):
        from .device import DeviceMesh
        if getattr(self, '_block', None) is None:
            m0, m1 = self.element_block()
            self._block = DeviceMesh(self.local.coords, self.local.elements[m0:m1],
                                     device=self.local.device, n_nodes=self.n_nodes, check=False)
        return self._block

    def _block_values(self, values, lead):
        """values for all M elements, or for the rank's block only -> the block's"""
        m0, m1 = self.element_block()
        m = self.local.n_elements
        if values.shape[-1] == m and m != m1 - m0:
            values = values[..., m0:m1]
        if values.shape[-1] != m1 - m0:
            raise ValueError('expected values for %d (all) or %d (the block\'s) elements, got %s'
                             % (m, m1 - m0, tuple(values.shape)))
        return values.contiguous()

    def load_vector_reduced(self, fq, weights, points):
        """The cubature load vector (solvers.py:540-598), complete on every rank.
        fq: (n_qp, M) values of f at the quadrature points of all elements, or
        (n_qp, m1 - m0) for the rank's element_block() only."""
        from .device import to_device_f64
        fq = self._block_values(to_device_f64(fq, self.local.device), 0)
        b = self._block_mesh().load_vector(fq, weights, points)
        dist.all_reduce(b, group=self.group)
        return b

    def load_vector_midpoint_reduced(self, f_centroid):
        """The legacy load vector (solvers.py:414-426), complete on every rank.
        f_centroid: (M,) or the block's (m1 - m0,) values."""
        from .device import to_device_f64
        fc = self._block_values(to_device_f64(f_centroid, self.local.device), 0)
        b = self._block_mesh().load_vector_midpoint(fc)
        dist.all_reduce(b, group=self.group)
        return b

    def element_block(self):
        """The calling rank's contiguous block of elements [m0, m1) (the estimator is
        per element: it is dealt out by elements, not by rows)."""
        m = self.local.n_elements
        return m * self.rank // self.world, m * (self.rank + 1) // self.world

    def eta_r(self, x, dirichlet, neumann, neumann_term, f_centroid):
        """compute_eta_r (indicators.py:7-86) of the rank's element block: (m1 - m0,)
        indicators, equal to the one-GPU result bit for bit.  Arguments as
        DeviceMesh.eta_r (global boundary lists, f at all centroids).  No collective:
        concatenating the ranks' pieces in rank order gives the whole vector."""
        m0, m1 = self.element_block()
        return self.local.eta_r_block(x, dirichlet, neumann, neumann_term, f_centroid, m0, m1)

    def global_chunk_offsets(self, indptr):
        return global_chunk_offsets(indptr, self.n_nodes, self.world, self.rank,
                                    self.chunk_log2, group=self.group)

    def global_row_pointer(self, indptr):
        return global_row_pointer(indptr, self.n_nodes, self.world, self.rank, self.chunk_log2,
                                  group=self.group)

    def global_rows(self):
        return self.local.global_rows()

    def gather(self, indptr, indices, data, root=0):
        return gather_csr(indptr, indices, data, self.global_rows(), self.n_nodes,
                          root=root, group=self.group)

    def close(self):
        if getattr(self, '_block', None) is not None:
            self._block.close()
            self._block = None
        self.local.close()
