"""Drop-in for ``p1afempy.refinement.refineNVB`` (with ``generate_new_nodes``,
/root/reference/p1afempy/refinement.py:8-31, 508-718) -- SURVEY.md section 8(f),
"next" rank 3: the step right before the assembly path.

Everything runs on the device: the edge numbering is the assembly path's
``provide_geometric_data`` (p1_geometric_data), marking with closure, numbering of
the new nodes, midpoints / embedded values, refined boundaries and children are the
CUDA kernels of csrc/afem.cu (p1_nvb_plan, p1_nvb_apply).  The refined mesh is the
reference's bit for bit: same node numbers, same element order.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import DeviceMesh, _ptr, _stream, to_device_f64, to_device_indices, to_host

__all__ = ['refineNVB', 'refine_nvb_device']


def refine_nvb_device(coords, elements, marked, boundaries, to_embed=None,
                      sort_for_longest_edge=False):
    """Newest vertex bisection on CUDA tensors: coords (N, 2) float64, elements (M, 3)
    int32 (rotated IN PLACE when sort_for_longest_edge, like the reference), marked:
    int64 element indices, boundaries: list of (K_j, 2) int32, to_embed: (N,) float64
    or None.  Returns (coords, elements, boundaries, embedded or None)."""
    dev = coords.device
    n, m = int(coords.shape[0]), int(elements.shape[0])
    with DeviceMesh(None, elements, device=dev.index, n_nodes=n, check=False) as mesh:
        lib, ctx, stream = mesh._lib, mesh.ctx, _stream(mesh.device)
        if sort_for_longest_edge and m:
            _lib.check(lib.p1_nvb_sort_longest_edge(ctx, _ptr(coords), _ptr(elements), m, stream))
        e2e, e2n, b2e = mesh.geometric_data(boundaries)
    n_edges = int(e2n.shape[0])
    offsets = np.zeros(len(boundaries) + 1, dtype=np.int64)
    for j, b in enumerate(boundaries):
        offsets[j + 1] = offsets[j] + int(b.shape[0])
    k = int(offsets[-1])
    bnd = (torch.cat([b.reshape(-1, 2) for b in boundaries]).contiguous() if k
           else torch.zeros((0, 2), dtype=torch.int32, device=dev))
    b2e_all = (torch.cat(list(b2e)).contiguous() if k
               else torch.zeros(0, dtype=torch.int64, device=dev))
    marked = marked.to(device=dev, dtype=torch.int64).contiguous()
    i32 = dict(dtype=torch.int32, device=dev)
    edge_flag = torch.empty(n_edges + 1, **i32)
    edge_new = torch.empty(n_edges + 1, **i32)
    child_first = torch.empty(m + 1, **i32)
    bnd_before = torch.empty(k + 1, **i32)
    counts = np.zeros(2 + len(boundaries), dtype=np.int64)
    offs_p = offsets.ctypes.data_as(C.c_void_p)
    _lib.check(lib.p1_nvb_plan(ctx, m, n_edges, _ptr(e2e), _ptr(marked), marked.numel(),
                               _ptr(b2e_all), offs_p, len(boundaries), _ptr(edge_flag),
                               _ptr(edge_new), _ptr(child_first), _ptr(bnd_before),
                               counts.ctypes.data_as(C.c_void_p), stream))
    n_new, m_new = int(counts[0]), int(counts[1])
    new_coords = torch.empty((n + n_new, 2), dtype=torch.float64, device=dev)
    new_coords[:n] = coords
    new_embed = None
    if to_embed is not None:
        new_embed = torch.empty(n + n_new, dtype=torch.float64, device=dev)
        new_embed[:n] = to_embed
    new_elements = torch.empty((m_new, 3), **i32)
    new_bnd = [torch.empty((int(b.shape[0]) + int(counts[2 + j]), 2), **i32)
               for j, b in enumerate(boundaries)]
    ptrs = (C.c_void_p * max(1, len(new_bnd)))(*[t.data_ptr() for t in new_bnd])
    _lib.check(lib.p1_nvb_apply(
        ctx, _ptr(coords), n, _ptr(elements), m, n_edges, _ptr(e2e), _ptr(e2n), _ptr(bnd),
        _ptr(b2e_all), offs_p, len(boundaries), _ptr(edge_flag), _ptr(edge_new),
        _ptr(child_first), _ptr(bnd_before), counts.ctypes.data_as(C.c_void_p),
        _ptr(to_embed), 1 if sort_for_longest_edge else 0, _ptr(new_coords), _ptr(new_embed),
        _ptr(new_elements), C.cast(ptrs, C.c_void_p), stream))
    return new_coords, new_elements, new_bnd, new_embed


def refineNVB(coordinates, elements, marked_elements, boundary_conditions,
              to_embed=np.array([]), sort_for_longest_egde=False):
    """Newest vertex bisection of the marked elements (refinement.py:508-718): returns
    (new_coordinates, new_elements, new_boundaries, embedded_values) like the
    reference -- elements / boundaries int64, ``to_embed`` interpolated onto the new
    nodes (returned unchanged when empty).  ``sort_for_longest_egde=True`` rotates
    ``elements`` in place first, as the reference does."""
    from .device import context
    _, device = context(None)
    dev = torch.device('cuda', device)
    ctx = context(device)[0]
    ct = to_device_f64(coordinates, device)
    et = to_device_indices(elements, device, ctx)
    if et.numel() == 0:
        et = et.reshape(0, 3)
    bts = [to_device_indices(np.asarray(b).reshape(-1, 2), device, ctx).reshape(-1, 2)
           for b in boundary_conditions]
    embed = np.asarray(to_embed)
    emb_t = to_device_f64(embed, device) if embed.size else None
    marked = torch.from_numpy(np.ascontiguousarray(np.asarray(marked_elements, dtype=np.int64)
                                                   .reshape(-1))).to(dev)
    nc, ne, nb, nemb = refine_nvb_device(ct, et, marked, bts, emb_t, bool(sort_for_longest_egde))
    if sort_for_longest_egde and isinstance(elements, np.ndarray) and elements.size:
        elements[...] = to_host(et).astype(elements.dtype)      # the reference's side effect
    out_b = []
    for b_in, b_new in zip(boundary_conditions, nb):
        b_in = np.asarray(b_in)
        out_b.append(to_host(b_new).astype(np.int64) if b_in.size else b_in)
    return (to_host(nc), to_host(ne).astype(np.int64), out_b,
            to_host(nemb) if nemb is not None else to_embed)
