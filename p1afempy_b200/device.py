"""Device-resident side of the drop-in: buffers (torch), the plan, and one
method per C-ABI entry point.  PyTorch is used for device memory and streams
only; all arithmetic happens in libp1b200.so.

``DeviceMesh`` is also the public device-resident API: torch CUDA tensors in,
CSR triple of torch CUDA tensors out (no PCIe traffic), for callers that keep
the mesh on the GPU across an adaptive loop.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib, options

_contexts = {}
_ctx_lock = threading.Lock()


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.P1Error(
            'p1afempy_b200 needs a CUDA device (B200, sm_100a); there is no '
            'CPU fallback.')


def context(device: int | None = None):
    """Returns (ctx handle, device index); one context per device."""
    _require_cuda()
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    with _ctx_lock:
        if device not in _contexts:
            torch.cuda.init()
            with torch.cuda.device(device):
                torch.zeros(1, device='cuda')   # make sure a primary context exists
            handle = C.c_void_p()
            _lib.check(_lib.load().p1_ctx_create(device, C.byref(handle)))
            _contexts[device] = handle
        return _contexts[device], device


def trim_scratch(device: int | None = None, limit_bytes: int | None = None):
    """Returns the library's idle scratch blocks (cached between calls, invisible to
    torch's allocator) to the driver; limit_bytes also sets how much may stay cached
    from now on (default 48 GB)."""
    ctx, _ = context(device)
    lib = _lib.load()
    if limit_bytes is not None:
        _lib.check(lib.p1_ctx_set_cache_limit(ctx, int(limit_bytes)))
    _lib.check(lib.p1_ctx_trim(ctx))


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _host_f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


_STAGE_CHUNK = 32 << 20      # bytes per pinned staging buffer
_STAGE_WAVE = 6               # host copies in flight (threads): 43 GB/s on a 16-cpu host (4: 38, 8: 43; tools/upload_probe.py)
_stage = {}                   # device -> (buffers, events, side stream, thread pool)
_stage_lock = threading.Lock()


def upload(a: np.ndarray, device: int) -> torch.Tensor:
    """Host numpy array -> CUDA tensor.  A large PAGEABLE array (what a caller of the
    reference holds) is staged by hand: worker threads copy 32 MB pieces into pinned
    buffers (numpy's copy releases the GIL) while the previous pieces travel; the
    driver's own staging of a pageable cudaMemcpy moves ~10 GB/s, this ~3x that.
    Small arrays and pinned ones take torch's path."""
    dev = torch.device('cuda', device)
    a = np.ascontiguousarray(a)
    src_t = torch.from_numpy(a)
    if a.nbytes < 4 * _STAGE_CHUNK or src_t.is_pinned():
        return src_t.to(dev, non_blocking=False)
    from concurrent.futures import ThreadPoolExecutor
    with _stage_lock:
        if device not in _stage:
            with torch.cuda.device(device):
                bufs = [torch.empty(_STAGE_CHUNK, dtype=torch.uint8, pin_memory=True)
                        for _ in range(2 * _STAGE_WAVE)]
                _stage[device] = (bufs, [None] * len(bufs), torch.cuda.Stream(device=device),
                                  ThreadPoolExecutor(_STAGE_WAVE))
        bufs, events, side, pool = _stage[device]
        dst = torch.empty(a.shape, dtype=src_t.dtype, device=dev)
        src8 = a.reshape(-1).view(np.uint8)
        dst8 = dst.reshape(-1).view(torch.uint8)
        n = a.nbytes
        offsets = list(range(0, n, _STAGE_CHUNK))
        for w0 in range(0, len(offsets), _STAGE_WAVE):
            wave = offsets[w0:w0 + _STAGE_WAVE]
            half = (w0 // _STAGE_WAVE) % 2 * _STAGE_WAVE
            jobs = []
            for i, off in enumerate(wave):
                j = half + i
                if events[j] is not None:
                    events[j].synchronize()          # the buffer's previous piece has left
                size = min(_STAGE_CHUNK, n - off)
                jobs.append((j, off, size, pool.submit(
                    np.copyto, bufs[j].numpy()[:size], src8[off:off + size])))
            with torch.cuda.stream(side):
                for j, off, size, fut in jobs:
                    fut.result()
                    dst8[off:off + size].copy_(bufs[j][:size], non_blocking=True)
                    events[j] = torch.cuda.Event()
                    events[j].record(side)
        torch.cuda.current_stream(device).wait_stream(side)
        dst.record_stream(side)
    return dst


def to_device_f64(a, device, shape=None):
    """numpy / torch (any device) -> contiguous float64 CUDA tensor."""
    if isinstance(a, torch.Tensor):
        t = a.to(device=torch.device('cuda', device), dtype=torch.float64)
    else:
        t = upload(_host_f64(a), device)
    t = t.contiguous()
    if shape is not None:
        t = t.reshape(shape)
    return t


def to_device_indices(a, device, ctx):
    """Integer index array (numpy of any integer dtype, or torch) -> int32
    CUDA tensor of the same shape.  The narrowing runs on the device
    (p1_narrow_indices); values outside int32 become -1 and are reported by
    the range check of the consumer."""
    dev = torch.device('cuda', device)
    lib = _lib.load()
    if isinstance(a, torch.Tensor):
        if a.dtype == torch.int32:
            return a.to(dev).contiguous()
        src = a.to(device=dev, dtype=torch.int64).contiguous()
        kind = 0
        shape = tuple(src.shape)
    else:
        a = np.asarray(a)
        if a.dtype.kind not in 'iu':
            if a.size and not np.all(np.mod(a, 1) == 0):
                raise IndexError('arrays used as indices must be of integer type')
            a = a.astype(np.int64)
        shape = a.shape
        if a.dtype == np.int32:
            return upload(a, device)
        if a.dtype == np.uint32:
            src = upload(np.ascontiguousarray(a).view(np.int32), device)
            kind = 1
        else:
            if a.dtype != np.int64:
                a = a.astype(np.int64)
            src = upload(a, device)
            kind = 0
    dst = torch.empty(shape, dtype=torch.int32, device=dev)
    _lib.check(lib.p1_narrow_indices(ctx, _ptr(src), kind, src.numel(), _ptr(dst),
                                     _stream(device)))
    return dst


def to_host(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> numpy through a pinned staging buffer (torch caches
    pinned blocks, so repeated calls do not re-register memory)."""
    if t.numel() == 0:
        return np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=False)
    return host.numpy()


def interleaved_count(n_rows, world, rank, chunk_log2):
    """Rows rank `rank` owns when chunks of 2^chunk_log2 rows are dealt round-robin
    (plan.cuh: Own::count)."""
    c = 1 << chunk_log2
    full, tail = divmod(n_rows, c)
    mine = (full // world) * c + (c if (full % world) > rank else 0)
    if tail and (full % world) == rank:
        mine += tail
    return mine


def scipy_index_dtype(n_elements, n_rows, n_cols):
    """Index dtype scipy gives ``coo_matrix(...).tocsr()`` for 9M triplets
    (scipy/sparse/_coo.py:419: maxval = max(nnz, N))."""
    big = max(9 * n_elements, n_rows, n_cols)
    return np.int64 if big > np.iinfo(np.int32).max else np.int32


class ColdAssembly:
    """What ``DeviceMesh.assemble_cold`` returns: the CSR arrays of a cold assembly
    that is still in flight.  indices / data have a capacity >= nnz, indptr[-1]
    (device) is nnz; ``wait()`` synchronises, checks the status words and returns
    exact-size views -- or the plan route's result when the mesh needed it (a node
    with more than 8 elements, a non-manifold vertex, nnz above the capacity)."""

    def __init__(self, mesh, indptr, indices, data, status, redo):
        self.mesh, self.indptr, self.indices, self.data = mesh, indptr, indices, data
        self.status, self._redo, self._done = status, redo, None

    def wait(self):
        if self._done is None:
            st = self.status.cpu().numpy()          # synchronises with the producing stream
            if st[0] & _lib.STATUS_INDEX:
                raise IndexError('element index out of range [0, %d)' % self.mesh.n_nodes)
            if st[0] & (_lib.STATUS_SLOWPATH | _lib.STATUS_RETRY):
                self._done = self._redo()
            else:
                nnz = int(st[6])
                self.mesh.nnz, self.mesh.max_index = nnz, int(st[1])
                self._done = (self.indptr,
                              None if self.indices is None else self.indices[:nnz],
                              self.data[:nnz])
            self._redo = None
        return self._done


class DeviceMesh:
    """Coordinates + elements on one GPU, plus the plan derived from
    ``elements`` (node -> incident elements, CSR row pointer).

    rows=(begin, end) restricts the plan to a contiguous row range (one rank of
    a multi-GPU assembly); everything else is identical.  reuse=True keeps the
    plan across assemble() calls (new coefficients on the same mesh only
    recompute the values); the default rebuilds it, which is what the
    reference's stateless functions do.
    """

    def __init__(self, coordinates, elements, device=None, rows=None,
                 n_nodes=None, reuse=False, check=True, interleave=None):
        self.ctx, self.device = context(device)
        self._lib = _lib.load()
        self.elements = to_device_indices(elements, self.device, self.ctx)
        if self.elements.numel() == 0:
            self.elements = self.elements.reshape(0, 3)
        if self.elements.ndim != 2 or self.elements.shape[1] != 3:
            raise ValueError('elements must have shape (M, 3)')
        self.coords = None
        if coordinates is not None:
            self.coords = to_device_f64(coordinates, self.device)
            if self.coords.ndim != 2 or self.coords.shape[1] != 2:
                raise ValueError('coordinates must have shape (N, 2)')
            n_nodes = self.coords.shape[0]
        if n_nodes is None:
            n_nodes = int(self.elements.max().item()) + 1 if self.elements.numel() else 0
        self.n_nodes = int(n_nodes)
        self.n_elements = int(self.elements.shape[0])
        self.rows = (0, self.n_nodes) if rows is None else (int(rows[0]), int(rows[1]))
        # (world, rank, chunk_log2): own the row chunks with index % world == rank
        self.interleave = None if interleave is None else tuple(int(x) for x in interleave)
        if self.interleave is not None and rows is not None:
            raise ValueError('rows= and interleave= are mutually exclusive')
        self.n_rows = self.rows[1] - self.rows[0]     # owned rows; final once planned
        self.reuse = bool(reuse)
        self._plan = None
        self._plan_keeps_elements = False
        self._plan_topology = False
        self._topology_fits = True      # until the topology-only plan is refused
        self._direct_fits = True        # ... or the plan-free load vectors are
        self._cold_refill = True        # new values on a kept plan: through the cold path
        self._exact = False
        self.nnz = None
        self.max_index = None
        if check and self.elements.numel():
            # every kernel gathers by element index: reject bad input up front
            _lib.check(self._lib.p1_check_indices(
                self.ctx, _ptr(self.elements), self.elements.numel(), self.n_nodes,
                _stream(self.device)))

    def balanced_rows(self, world):
        """Row boundaries (world+1) with about equal numbers of stored entries
        per range; deterministic, identical on every rank (p1_partition_rows)."""
        out = np.zeros(world + 1, dtype=np.int64)
        _lib.check(self._lib.p1_partition_rows(
            self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes, int(world),
            out.ctypes.data_as(C.c_void_p), _stream(self.device)))
        return [int(x) for x in out]

    def set_rows(self, rows):
        self.close()
        self.rows = (int(rows[0]), int(rows[1]))

    # ------------------------------------------------------------------ plan
    def _flags(self, keep):
        return ((_lib.PLAN_EXACT if self._exact else 0)
                | (_lib.PLAN_KEEP_ELEMENTS if keep else 0)
                | int(options.plan_flags))

    def _adopt(self, handle, keep, topology=False):
        self._plan = handle
        self._plan_keeps_elements = keep
        self._plan_topology = topology
        self.n_rows = int(self._lib.p1_plan_rows(handle))
        nnz, mx = C.c_int64(), C.c_int64()
        _lib.check(self._lib.p1_plan_info(handle, C.byref(nnz), C.byref(mx), None, None))
        self.nnz, self.max_index = nnz.value, mx.value

    def _create(self, keep, op, local):
        handle = C.c_void_p()
        if self.interleave is not None:
            world, rank, chunk_log2 = self.interleave
            rc = self._lib.p1_plan_create_interleaved(
                self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes, chunk_log2,
                world, rank, self._flags(keep), op, _ptr(self.coords) if op else None,
                _ptr(local), _stream(self.device), C.byref(handle))
        elif op:
            rc = self._lib.p1_plan_create_for(
                self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes,
                self.rows[0], self.rows[1], self._flags(keep), op, _ptr(self.coords),
                _ptr(local), _stream(self.device), C.byref(handle))
        else:
            rc = self._lib.p1_plan_create(
                self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes,
                self.rows[0], self.rows[1], self._flags(keep), _stream(self.device),
                C.byref(handle))
        _lib.check(rc)
        self._adopt(handle, keep)

    @property
    def plan(self):
        """Symbolic plan that remembers the element of every record (what load
        vectors, edge numbering and eta_R need)."""
        return self._element_plan(False)

    def _element_plan(self, all_twins):
        """all_twins: the caller (eta_R) needs the twin of EVERY half-edge; a fresh
        topology-only plan then also records the slot of every half-edge
        (P1_PLAN_SLOT_INDEX)."""
        if self._plan is not None and not (self._plan_keeps_elements or self._plan_topology):
            self.close()
        if self._plan is None and not self.reuse and self._topology_fits \
                and self.interleave is None:
            # no matrix will be refreshed from this plan: 16 bytes per incidence
            # instead of 36, no row pointer (P1_PLAN_TOPOLOGY)
            handle = C.c_void_p()
            rc = self._lib.p1_plan_create(
                self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes,
                self.rows[0], self.rows[1],
                _lib.PLAN_TOPOLOGY | (_lib.PLAN_SLOT_INDEX if all_twins else 0),
                _stream(self.device), C.byref(handle))
            if rc == _lib.P1_ERETRY:
                self._topology_fits = False     # a node with more than 8 elements
            else:
                _lib.check(rc)
                self._adopt(handle, False, True)
        if self._plan is None:
            self._create(True, 0, None)
        return self._plan

    def _plan_for(self, op, local):
        """Cold path of one matrix: scatter and local matrices in one kernel."""
        self._create(self.reuse, op, local)

    def _assemble_quad(self, op, args, keepalive, pattern):
        """Cold generalised stiffness / weighted mass: quadrature sums and local
        matrices inside the scatter (p1_plan_create_quad), then P1_OP_STORED."""
        for attempt in (0, 1):
            self.close()
            handle = C.c_void_p()
            _lib.check(self._lib.p1_plan_create_quad(
                self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes,
                self.rows[0], self.rows[1], self._flags(False), op, _ptr(self.coords),
                C.byref(args), _stream(self.device), C.byref(handle)))
            self._adopt(handle, False)
            data = self._empty(self.nnz)
            indptr = self._empty(self.n_rows + 1, torch.int32) if pattern else None
            indices = self._empty(self.nnz, torch.int32) if pattern else None
            rc = self._lib.p1_assemble_csr(self._plan, _lib.OP_STORED, None, None,
                                           _ptr(indptr), _ptr(indices), _ptr(data),
                                           _stream(self.device))
            if rc == _lib.P1_ERETRY and attempt == 0 and not self._exact:
                self._exact = True
                continue
            _lib.check(rc)
            self.close()
            del keepalive
            return indptr, indices, data

    def global_rows(self):
        """Global row index of every local row (int64 CUDA tensor)."""
        dev = torch.device('cuda', self.device)
        if self.interleave is None:
            return torch.arange(self.rows[0], self.rows[1], dtype=torch.int64, device=dev)
        world, rank, c = self.interleave
        n = self.n_nodes
        chunks = torch.arange(rank, max(rank, (n + (1 << c) - 1) >> c), world, dtype=torch.int64,
                              device=dev)
        rows = (chunks[:, None] << c) + torch.arange(1 << c, dtype=torch.int64, device=dev)
        rows = rows.reshape(-1)
        return rows[rows < n]

    def close(self):
        if self._plan is not None:
            self._lib.p1_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _empty(self, shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=torch.device('cuda', self.device))

    # -------------------------------------------------------------- assembly
    def assemble(self, op, local=None, pattern=True):
        """-> (indptr, indices, data) CUDA tensors over the owned rows.
        pattern=False skips the index arrays (re-assembly on the same mesh)."""
        if (self._plan is not None and self._plan_keeps_elements and self._cold_refill
                and self.interleave is None and self.rows == (0, self.n_nodes)):
            # New values on an unchanged mesh.  Refreshing the plan's records in place and
            # filling from them (p1_assemble_csr with an operator: one thread per record
            # slot recomputes its element) takes 6.9 ms at 64M triangles; the whole cold
            # assembly without the index arrays takes 4.2 and gives the same bits.
            cold = self.assemble_cold(op, local=local, pattern=pattern)
            if int(cold.status.cpu()[0]) == 0:
                indptr, indices, data = cold.wait()
                return (indptr if pattern else None), indices, data
            self._cold_refill = False            # this mesh needs the plan route
        for attempt in (0, 1):
            if self._plan is not None and self._plan_keeps_elements:
                run_op = op                      # refresh the values in place
            else:
                self.close()
                self._plan_for(op, local)
                run_op = _lib.OP_STORED
            n_own = self.n_rows
            data = self._empty(self.nnz)
            indptr = self._empty(n_own + 1, torch.int32) if pattern else None
            indices = self._empty(self.nnz, torch.int32) if pattern else None
            rc = self._lib.p1_assemble_csr(
                self._plan, run_op, _ptr(self.coords), _ptr(local), _ptr(indptr),
                _ptr(indices), _ptr(data), _stream(self.device))
            if rc == _lib.P1_ERETRY and attempt == 0 and not self._exact:
                # non-manifold vertex / degenerate element: the closed-fan
                # shortcut of the default plan does not hold; count exactly
                self.close()
                self._exact = True
                continue
            _lib.check(rc)
            if not self.reuse:
                self.close()
            return indptr, indices, data

    def cold_capacity(self):
        """Entries the outputs of ``assemble_cold`` hold: rows + incidences + one per
        row for an open fan (all rows owned: a bound; a share of the rows: the
        expected count plus 30 %)."""
        if self.interleave is None and self.rows == (0, self.n_nodes):
            return 2 * self.n_nodes + 3 * self.n_elements
        world = self.interleave[0] if self.interleave is not None else max(
            1, self.n_nodes // max(1, self.rows[1] - self.rows[0]))
        return 2 * self.n_rows + int(1.3 * 3 * self.n_elements / world) + 4096

    def assemble_cold(self, op, local=None, pattern=True, quad=None, keepalive=None, out=None):
        """Cold assembly without a host synchronisation (p1_assemble_cold): returns a
        ColdAssembly at once; nothing is cached on the mesh.  op: OP_STIFFNESS /
        OP_MASS / OP_LOCAL (local: (M, 9) CUDA tensor) / OP_GENERAL / OP_WEIGHTED
        (quad: _lib.QuadArgs).  out: a ColdAssembly of an earlier call on a mesh of
        the same size, whose arrays are overwritten (a time-stepping loop; on meshes of
        at most 4M triangles the library then replays the call as one CUDA graph)."""
        if self.interleave is not None:
            world, rank, chunk_log2 = self.interleave
        elif self.rows == (0, self.n_nodes):
            world, rank, chunk_log2 = 1, 0, 0
        else:
            raise ValueError('assemble_cold: all rows or interleave=')
        if self.interleave is not None:
            self.n_rows = int(interleaved_count(self.n_nodes, world, rank, chunk_log2))
        cap = self.cold_capacity()
        if out is not None:
            indptr, indices, data, status = out.indptr, out.indices, out.data, out.status
            if (indptr.numel() != self.n_rows + 1 or data.numel() != cap
                    or (indices is None) != (not pattern) or data.device.index != self.device):
                raise ValueError('assemble_cold: out= belongs to a mesh of another size')
        else:
            indptr = self._empty(self.n_rows + 1, torch.int32)
            indices = self._empty(cap, torch.int32) if pattern else None
            data = self._empty(cap)
            status = self._empty(8, torch.int32)
        _lib.check(self._lib.p1_assemble_cold(
            self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes, chunk_log2, world,
            rank, op, _ptr(self.coords), _ptr(local), C.byref(quad) if quad is not None else None,
            cap, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(status), _stream(self.device)))

        def redo(keep=(local, quad, keepalive)):
            if op in (_lib.OP_GENERAL, _lib.OP_WEIGHTED):
                return self._assemble_quad(op, keep[1], keep[2], pattern)
            return self.assemble(op, local=keep[0], pattern=pattern)
        return ColdAssembly(self, indptr, indices, data, status, redo)

    def stiffness_csr(self, pattern=True):
        return self.assemble(_lib.OP_STIFFNESS, pattern=pattern)

    def mass_csr(self, pattern=True):
        return self.assemble(_lib.OP_MASS, pattern=pattern)

    def general_stiffness_csr(self, a11q, a12q, a21q, a22q, weights, pattern=True,
                              fused=True):
        """a_ij_q: (n_qp, M) CUDA float64 values of the coefficients at the
        quadrature points (see ``quadrature_points``).  fused: cold calls evaluate
        the quadrature sums and the local matrix inside the scatter
        (p1_plan_create_quad) instead of writing an [M x 9] array first."""
        w = _host_f64(weights)
        qs = [to_device_f64(a, self.device, (w.shape[0], self.n_elements))
              for a in (a11q, a12q, a21q, a22q)]
        if fused and not self.reuse and self.interleave is None:
            # 64M triangles, 16 points: 10.6 ms, against 12.4 ms through
            # p1_local_general_stiffness + P1_OP_LOCAL (since the scatter keeps four
            # points of the four arrays in flight; before that the two-step route won)
            args = _lib.QuadArgs(qs[0].data_ptr(), qs[1].data_ptr(), qs[2].data_ptr(),
                                 qs[3].data_ptr(), None, w.ctypes.data, None, w.shape[0])
            return self._assemble_quad(_lib.OP_GENERAL, args, (qs, w), pattern)
        local = self._empty((self.n_elements, 9))
        _lib.check(self._lib.p1_local_general_stiffness(
            self.ctx, _ptr(self.coords), _ptr(self.elements), self.n_elements,
            _ptr(qs[0]), _ptr(qs[1]), _ptr(qs[2]), _ptr(qs[3]),
            w.ctypes.data_as(C.c_void_p), w.shape[0], _ptr(local),
            _stream(self.device)))
        return self.assemble(_lib.OP_LOCAL, local=local, pattern=pattern)

    def weighted_mass_csr(self, phiq, weights, points, pattern=True):
        w, p = _host_f64(weights), _host_f64(points)
        phiq = to_device_f64(phiq, self.device, (w.shape[0], self.n_elements))
        if not self.reuse and self.interleave is None:
            args = _lib.QuadArgs(None, None, None, None, phiq.data_ptr(), w.ctypes.data,
                                 p.ctypes.data, w.shape[0])
            return self._assemble_quad(_lib.OP_WEIGHTED, args, (phiq, w, p), pattern)
        local = self._empty((self.n_elements, 9))
        _lib.check(self._lib.p1_local_weighted_mass(
            self.ctx, _ptr(self.coords), _ptr(self.elements), self.n_elements,
            _ptr(phiq), w.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
            w.shape[0], _ptr(local), _stream(self.device)))
        return self.assemble(_lib.OP_LOCAL, local=local, pattern=pattern)

    def triplets(self, stiffness=False):
        m = self.n_elements
        rows = self._empty(9 * m, torch.int64)
        cols = self._empty(9 * m, torch.int64)
        vals = self._empty(9 * m)
        fn = self._lib.p1_stiffness_triplets if stiffness else self._lib.p1_mass_triplets
        _lib.check(fn(self.ctx, _ptr(self.coords), _ptr(self.elements), m, _ptr(rows),
                      _ptr(cols), _ptr(vals), _stream(self.device)))
        return rows, cols, vals

    # ----------------------------------------------------- per-element data
    def geometry(self, area=True, vectors=False):
        m = self.n_elements
        a = self._empty(m) if area else None
        d21 = self._empty((m, 2)) if vectors else None
        d31 = self._empty((m, 2)) if vectors else None
        _lib.check(self._lib.p1_geometry(self.ctx, _ptr(self.coords), _ptr(self.elements),
                                         m, _ptr(a), _ptr(d21), _ptr(d31),
                                         _stream(self.device)))
        return a, d21, d31

    def quadrature_points(self, points):
        p = _host_f64(points).reshape(-1, 2)
        out = self._empty((p.shape[0], self.n_elements, 2))
        _lib.check(self._lib.p1_quadrature_points(
            self.ctx, _ptr(self.coords), _ptr(self.elements), self.n_elements,
            p.ctypes.data_as(C.c_void_p), p.shape[0], _ptr(out), _stream(self.device)))
        return out

    def nodal_at_quadrature(self, u, points):
        p = _host_f64(points).reshape(-1, 2)
        u = to_device_f64(u, self.device, (self.n_nodes,))
        out = self._empty((p.shape[0], self.n_elements))
        _lib.check(self._lib.p1_nodal_at_quadrature(
            self.ctx, _ptr(u), _ptr(self.elements), self.n_elements,
            p.ctypes.data_as(C.c_void_p), p.shape[0], _ptr(out), _stream(self.device)))
        return out

    def centroids(self):
        out = self._empty((self.n_elements, 2))
        _lib.check(self._lib.p1_centroids(self.ctx, _ptr(self.coords), _ptr(self.elements),
                                          self.n_elements, _ptr(out),
                                          _stream(self.device)))
        return out

    def boundary_edges(self, edges_i32):
        k = int(edges_i32.shape[0])
        length = self._empty(k)
        mid = self._empty((k, 2))
        _lib.check(self._lib.p1_boundary_edges(self.ctx, _ptr(self.coords),
                                               _ptr(edges_i32), k, _ptr(length),
                                               _ptr(mid), _stream(self.device)))
        return length, mid

    # ---------------------------------------------------------- load vectors
    def _direct_vectors(self):
        """No plan to reuse and none wanted: load vectors go through the plan-free
        entry points (p1_load_vector*_direct)."""
        return (self._plan is None and not self.reuse and self._direct_fits
                and self.interleave is None and self.rows == (0, self.n_nodes))

    def load_vector(self, fq, weights, points):
        w, p = _host_f64(weights), _host_f64(points)
        fq = to_device_f64(fq, self.device, (w.shape[0], self.n_elements))
        if self._direct_vectors():
            b = self._empty(self.n_nodes)
            rc = self._lib.p1_load_vector_direct(
                self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes, _ptr(self.coords),
                _ptr(fq), w.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                w.shape[0], _ptr(b), _stream(self.device))
            if rc != _lib.P1_ERETRY:
                _lib.check(rc)
                return b
            self._direct_fits = False           # too many nodes with more than 8 elements
        self.plan  # noqa: B018  (fixes n_rows)
        b = self._empty(self.n_rows)
        _lib.check(self._lib.p1_load_vector(
            self.plan, _ptr(self.coords), _ptr(fq), w.ctypes.data_as(C.c_void_p),
            p.ctypes.data_as(C.c_void_p), w.shape[0], _ptr(b), _stream(self.device)))
        return b

    def integrate(self, fq, weights):
        """sum_T sum_q w_q fq[q][T] 2|T| (host float)."""
        w = _host_f64(weights)
        fq = to_device_f64(fq, self.device, (w.shape[0], self.n_elements))
        out = C.c_double()
        _lib.check(self._lib.p1_integrate_quadrature(
            self.ctx, _ptr(self.coords), _ptr(self.elements), self.n_elements, _ptr(fq),
            w.ctypes.data_as(C.c_void_p), w.shape[0], C.byref(out), _stream(self.device)))
        return out.value

    def load_vector_midpoint(self, f_centroid):
        fc = to_device_f64(f_centroid, self.device, (self.n_elements,))
        if self._direct_vectors():
            b = self._empty(self.n_nodes)
            rc = self._lib.p1_load_vector_midpoint_direct(
                self.ctx, _ptr(self.elements), self.n_elements, self.n_nodes, _ptr(self.coords),
                _ptr(fc), _ptr(b), _stream(self.device))
            if rc != _lib.P1_ERETRY:
                _lib.check(rc)
                return b
            self._direct_fits = False
        self.plan  # noqa: B018
        b = self._empty(self.n_rows)
        _lib.check(self._lib.p1_load_vector_midpoint(
            self.plan, _ptr(self.coords), _ptr(fc), _ptr(b), _stream(self.device)))
        return b

    # ------------------------------------------------- edges and estimator
    def geometric_data(self, boundaries):
        """boundaries: list of (K_j, 2) integer arrays (host or device).
        -> (element2edges (M,3), edge2nodes (E,2), [K_j] edge ids), int64 CUDA."""
        dev = torch.device('cuda', self.device)
        parts = [to_device_indices(np.asarray(b).reshape(-1, 2)
                                   if not isinstance(b, torch.Tensor) else b,
                                   self.device, self.ctx).reshape(-1, 2)
                 for b in boundaries]
        offsets = np.zeros(len(parts) + 1, dtype=np.int64)
        for j, b in enumerate(parts):
            offsets[j + 1] = offsets[j] + b.shape[0]
        k = int(offsets[-1])
        bnd = (torch.cat(parts).contiguous() if parts
               else torch.empty((0, 2), dtype=torch.int32, device=dev))
        m = self.n_elements
        e2e = self._empty((m, 3), torch.int64)
        cap = (3 * m + k + 1) // 2
        e2n = self._empty((cap, 2), torch.int64)
        b2e = self._empty(k, torch.int64)
        n_edges = C.c_int64()
        _lib.check(self._lib.p1_geometric_data(
            self.plan, _ptr(bnd), offsets.ctypes.data_as(C.c_void_p), len(parts),
            _ptr(e2e), _ptr(e2n), _ptr(b2e), C.byref(n_edges), _stream(self.device)))
        pieces = [b2e[offsets[j]:offsets[j + 1]] for j in range(len(parts))]
        return e2e, e2n[:n_edges.value], pieces

    def element_neighbours(self):
        """(M, 3) int64: the element across every edge, -1 on the boundary
        (mesh.py:495-528)."""
        out = self._empty((self.n_elements, 3), torch.int64)
        _lib.check(self._lib.p1_element_neighbours(self._element_plan(True), _ptr(out),
                                                   _stream(self.device)))
        return out

    def eta_r(self, x, dirichlet, neumann, neumann_term, f_centroid, partial=False):
        """partial: this mesh is a cut-out of a larger one (eta_r_block): open edges
        that no boundary lists are the cut, not an error."""
        x = to_device_f64(x, self.device, (self.n_nodes,))
        fc = to_device_f64(f_centroid, self.device, (self.n_elements,))
        nt = (to_device_f64(neumann_term, self.device, (neumann.shape[0],))
              if neumann.shape[0] else None)
        eta = self._empty(self.n_elements)
        fn = self._lib.p1_eta_r_partial if partial else self._lib.p1_eta_r
        _lib.check(fn(
            self._element_plan(True), _ptr(self.coords), _ptr(x),
            _ptr(dirichlet) if dirichlet.shape[0] else None, int(dirichlet.shape[0]),
            _ptr(neumann) if neumann.shape[0] else None, int(neumann.shape[0]),
            _ptr(nt), _ptr(fc), _ptr(eta), _stream(self.device)))
        return eta

    def eta_r_block(self, x, dirichlet, neumann, neumann_term, f_centroid, m0, m1):
        """eta_R of the elements [m0, m1) only -- one rank's share of a multi-GPU
        estimator (indicators.py:7-86): computed on the sub-mesh of the elements that
        touch a node of the block (p1_block_submesh), where the block's indicators are
        the global ones bit for bit.  dirichlet / neumann / neumann_term: the GLOBAL
        boundary lists; f_centroid: (M,) values for all elements."""
        m, n = self.n_elements, self.n_nodes
        m0, m1 = int(m0), int(m1)
        flag = self._empty(n, torch.uint8)
        pos = self._empty(m + 1, torch.int32)
        counts = np.zeros(2, dtype=np.int64)
        _lib.check(self._lib.p1_block_submesh(
            self.ctx, _ptr(self.elements), m, n, m0, m1, _ptr(flag), _ptr(pos),
            counts.ctypes.data_as(C.c_void_p), _stream(self.device)))
        n_sub, start = int(counts[0]), int(counts[1])
        sub = self._empty((n_sub, 3), torch.int32)
        ids = self._empty(n_sub, torch.int32)
        _lib.check(self._lib.p1_block_submesh_fill(
            self.ctx, _ptr(self.elements), m, _ptr(pos), _ptr(sub), _ptr(ids),
            _stream(self.device)))
        del pos

        def inside(edges):      # boundary edges of the sub-mesh that matter: both nodes flagged
            if edges.shape[0] == 0:
                return torch.zeros(0, dtype=torch.bool, device=flag.device)
            return (flag[edges[:, 0].long()] & flag[edges[:, 1].long()]).bool()
        kd, kn = inside(dirichlet), inside(neumann)
        d_sub, n_sub_edges = dirichlet[kd].contiguous(), neumann[kn].contiguous()
        term = None
        if neumann.shape[0]:
            term = to_device_f64(neumann_term, self.device, (neumann.shape[0],))[kn].contiguous()
        fc = to_device_f64(f_centroid, self.device, (m,))[ids.long()].contiguous()
        with DeviceMesh(self.coords, sub, device=self.device, check=False) as part:
            eta = part.eta_r(x, d_sub, n_sub_edges, term, fc, partial=True)
        return eta[start:start + (m1 - m0)]

    # ------------------------------------------------------------ host side
    def widen_indices(self, t):
        """int32 index tensor -> int64 (p1_widen_indices): what scipy's index dtype
        rule asks for when max(9 M, N) > 2^31 - 1."""
        t = t.contiguous()
        out = self._empty(t.shape, torch.int64)
        _lib.check(self._lib.p1_widen_indices(self.ctx, _ptr(t), t.numel(), _ptr(out),
                                              _stream(self.device)))
        return out

    def to_scipy(self, indptr, indices, data, inferred_shape):
        """CSR triple (CUDA) -> scipy.sparse.csr_matrix equal to
        ``reference(...).tocsr()``: shape max_index+1 when the reference lets
        scipy infer it (solvers.py:46-47,153) or (N, N) when it passes it
        (solvers.py:140-141,375-376); index dtype by scipy's rule."""
        from scipy.sparse import csr_matrix
        n = (self.max_index + 1) if inferred_shape else self.n_nodes
        if self.rows != (0, self.n_nodes) or self.interleave is not None:
            raise ValueError('to_scipy needs a plan over all rows')
        idx_dtype = scipy_index_dtype(self.n_elements, n, n)
        ip = indptr[:n + 1]
        if idx_dtype == np.int64:
            ip, indices = self.widen_indices(ip), self.widen_indices(indices)
        h_ip, h_ix, h_data = to_host(ip), to_host(indices), to_host(data)
        if n == 0:
            return csr_matrix((0, 0), dtype=np.float64)
        a = csr_matrix((h_data, h_ix, h_ip), shape=(n, n), copy=False)
        a.has_sorted_indices = True
        a.has_canonical_format = True
        return a
