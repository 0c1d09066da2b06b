"""Host-side mesh generators for tests and benchmarks (numpy only).

The hot path consumes meshes, it does not make them; the reference makes its
benchmark meshes with ``refineNVB`` (/root/reference/p1afempy/refinement.py:
508-718, driven by tests/manual/performance_test_stiffnes_matrix.py:15-31).
``/root/reference`` does not exist on the GPU box, so this module provides a
generator that yields THE SAME meshes (same node numbering, same element
order -- this matters, the memory access pattern of the kernels depends on
it); ``tests/test_meshgen_vs_reference.py`` checks that claim against the real
``refineNVB`` wherever the reference is importable.

Everything here is plain host code and deliberately simple; it is not part of
the accelerated path.
"""
from __future__ import annotations

import numpy as np


def number_edges(elements: np.ndarray, boundaries=()):
    """Edge numbering with the reference's convention (mesh.py:87-161):
    half-edge p = k*M+m runs e[m,k] -> e[m,(k+1)%3]; boundary edges are
    appended reversed; an undirected edge gets its number from the position
    of its (low -> high) occurrence.

    Returns (element2edges (M,3) int64, edge2nodes (E,2) int64,
    [boundary -> edge ids])."""
    elements = np.asarray(elements, dtype=np.int64)
    m = elements.shape[0]
    tails = [elements.T.reshape(-1)]
    heads = [elements[:, [1, 2, 0]].T.reshape(-1)]
    bounds = [3 * m]
    for b in boundaries:
        b = np.asarray(b, dtype=np.int64).reshape(-1, 2)
        tails.append(b[:, 1])
        heads.append(b[:, 0])
        bounds.append(bounds[-1] + b.shape[0])
    tail = np.concatenate(tails)
    head = np.concatenate(heads)
    span = int(max(tail.max(initial=0), head.max(initial=0))) + 1
    up = np.flatnonzero(tail < head)
    down = np.flatnonzero(head < tail)
    if up.size != down.size:
        raise ValueError('every boundary edge must be listed in a boundary')
    ids = np.empty(tail.size, dtype=np.int64)
    ids[up] = np.arange(up.size)
    key_up = tail[up] * span + head[up]
    key_down = head[down] * span + tail[down]
    ids[down[np.argsort(key_down, kind='stable')]] = \
        ids[up[np.argsort(key_up, kind='stable')]]
    e2e = ids[:3 * m].reshape(3, m).T.copy()
    e2n = np.stack([tail[up], head[up]], axis=1)
    b2e = [ids[bounds[i]:bounds[i + 1]] for i in range(len(bounds) - 1)]
    return e2e, e2n, b2e


def refine_nvb(coordinates, elements, marked, boundaries=()):
    """Newest vertex bisection; same output as the reference's ``refineNVB``
    (refinement.py:508-718, sort_for_longest_egde=False).

    Returns (coordinates, elements (int64), boundaries)."""
    coordinates = np.asarray(coordinates, dtype=np.float64)
    elements = np.asarray(elements, dtype=np.int64)
    boundaries = [np.asarray(b, dtype=np.int64).reshape(-1, 2)
                  for b in boundaries]
    m = elements.shape[0]
    n = coordinates.shape[0]
    e2e, e2n, b2e = number_edges(elements, boundaries)

    split = np.zeros(e2n.shape[0], dtype=bool)
    split[e2e[np.asarray(marked, dtype=np.int64)].reshape(-1)] = True
    while True:  # closure: a split edge forces the reference edge
        flags = split[e2e]
        need = ~flags[:, 0] & (flags[:, 1] | flags[:, 2])
        if not need.any():
            break
        split[e2e[need, 0]] = True

    new_id = np.zeros(e2n.shape[0], dtype=np.int64)
    chosen = np.flatnonzero(split)
    new_id[chosen] = n + np.arange(chosen.size)
    mid = (coordinates[e2n[chosen, 0]] + coordinates[e2n[chosen, 1]]) / 2.
    new_coordinates = np.vstack([coordinates, mid])

    new_boundaries = []
    for b, be in zip(boundaries, b2e):
        if b.size:
            nn = new_id[be]
            hit = np.flatnonzero(nn)
            if hit.size:
                b = np.vstack([b[nn == 0],
                               np.stack([b[hit, 0], nn[hit]], axis=1),
                               np.stack([nn[hit], b[hit, 1]], axis=1)])
        new_boundaries.append(b)

    nn = new_id[e2e]
    s0, s1, s2 = nn[:, 0] != 0, nn[:, 1] != 0, nn[:, 2] != 0
    kinds = {
        'b1': s0 & ~s1 & ~s2, 'b12': s0 & s1 & ~s2,
        'b13': s0 & ~s1 & s2, 'b123': s0 & s1 & s2}
    n_child = np.ones(m, dtype=np.int64)
    n_child[kinds['b1']] = 2
    n_child[kinds['b12']] = 3
    n_child[kinds['b13']] = 3
    n_child[kinds['b123']] = 4
    first = np.concatenate([[0], np.cumsum(n_child)])
    out = np.zeros((first[-1], 3), dtype=np.int64)
    first = first[:-1]
    keep = ~s0
    out[first[keep]] = elements[keep]

    v0, v1, v2 = elements[:, 0], elements[:, 1], elements[:, 2]
    m0, m1, m2 = nn[:, 0], nn[:, 1], nn[:, 2]

    def put(sel, children):
        base = first[sel]
        for j, (a, b, c) in enumerate(children):
            out[base + j, 0] = a[sel]
            out[base + j, 1] = b[sel]
            out[base + j, 2] = c[sel]

    put(kinds['b1'], [(v2, v0, m0), (v1, v2, m0)])
    put(kinds['b12'], [(v2, v0, m0), (m0, v1, m1), (v2, m0, m1)])
    put(kinds['b13'], [(m0, v2, m2), (v0, m0, m2), (v1, v2, m0)])
    put(kinds['b123'], [(m0, v2, m2), (v0, m0, m2), (m0, v1, m1),
                        (v2, m0, m1)])
    return new_coordinates, out, new_boundaries


def refine_uniform(coordinates, elements, boundaries=(), times=1):
    for _ in range(times):
        coordinates, elements, boundaries = refine_nvb(
            coordinates, elements, np.arange(elements.shape[0]), boundaries)
    return coordinates, elements, list(boundaries)


def criss_square():
    """S(0): unit square cut into 4 triangles by its centre
    (benchmark meshes)."""
    coordinates = np.array([[0., 0.], [1., 0.], [1., 1.], [0., 1.],
                            [.5, .5]])
    elements = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]],
                        dtype=np.int64)
    boundary = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.int64)
    return coordinates, elements, [boundary]


def unit_square(k: int):
    """S(k): ``criss_square`` refined k times with every element marked.
    M = 4^(k+1), N = 2n^2+2n+1 with n = 2^k."""
    c, e, b = criss_square()
    return refine_uniform(c, e, b, times=k)


def two_triangle_square():
    """The reference perf script's start mesh
    (performance_test_stiffnes_matrix.py:15-28)."""
    coordinates = np.array([[0., 0.], [1., 0.], [1., 1.], [0., 1.]])
    elements = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int64)
    boundary = np.array([[0, 1], [1, 2], [2, 3], [3, 0]], dtype=np.int64)
    return coordinates, elements, [boundary]


def l_shape():
    """6-triangle L-shaped domain, [0,2]^2 without (1,2]x(1,2], re-entrant
    corner at (1,1); the same mesh as the reference fixture
    tests/data/l_shape_mesh (8 nodes, 6 triangles, 3 boundary pieces)."""
    coordinates = np.array([[0., 0.], [1., 0.], [2., 0.], [2., 1.],
                            [1., 1.], [1., 2.], [0., 2.], [0., 1.]])
    elements = np.array([[0, 1, 4], [0, 4, 7], [1, 2, 4], [2, 3, 4],
                         [7, 4, 6], [4, 5, 6]], dtype=np.int64)
    boundaries = [np.array([[0, 1], [1, 2], [2, 3]], dtype=np.int64),
                  np.array([[3, 4], [4, 5], [5, 6]], dtype=np.int64),
                  np.array([[6, 7], [7, 0]], dtype=np.int64)]
    return coordinates, elements, boundaries


def perturb_interior(coordinates, boundaries, scale, seed=0):
    """Moves every node that is not on a listed boundary by a uniform random
    offset in [-scale, scale]^2 (dyadic meshes make the reference arithmetic
    exact; parity must also hold on generic coordinates)."""
    rng = np.random.default_rng(seed)
    fixed = np.zeros(coordinates.shape[0], dtype=bool)
    for b in boundaries:
        fixed[np.asarray(b).reshape(-1)] = True
    out = coordinates.copy()
    shift = rng.uniform(-scale, scale, size=coordinates.shape)
    out[~fixed] += shift[~fixed]
    return out


# ---------------------------------------------------------------------------
# device-side uniform refinement (benchmark meshes beyond what the host can hold)
# ---------------------------------------------------------------------------

def refine_uniform_device(coords, elements, boundaries):
    """One NVB step with EVERY element marked, on the GPU.

    coords (N,2) float64, elements (M,3) int32, boundaries: list of (K,2) int32
    CUDA tensors.  Same output as ``refine_nvb(c, e, arange(M), b)`` (hence as
    the reference's ``refineNVB``), bit for bit."""
    import torch
    marked = torch.arange(elements.shape[0], dtype=torch.int64, device=elements.device)
    return refine_nvb_device(coords, elements, marked, boundaries)


def unit_square_device(k: int, device=None, host_levels: int = 6):
    """S(k) as CUDA tensors (coords float64, elements int32, [boundary] int32):
    the first levels on the host, the rest refined on the device."""
    import torch
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    c, e, b = unit_square(min(k, host_levels))
    ct = torch.from_numpy(c).to(dev)
    et = torch.from_numpy(e.astype(np.int32)).to(dev)
    bt = [torch.from_numpy(x.astype(np.int32)).to(dev) for x in b]
    for _ in range(max(0, k - host_levels)):
        ct, et, bt = refine_uniform_device(ct, et, bt)
    return ct, et, bt


def refine_nvb_device(coords, elements, marked, boundaries):
    """Newest vertex bisection of the MARKED elements on the GPU; same output as
    ``refine_nvb`` / the reference's ``refineNVB`` (refinement.py:508-718), bit
    for bit.  coords (N,2) float64, elements (M,3) int32, marked int64 indices,
    boundaries: list of (K,2) int32 -- all CUDA tensors.  (CUDA kernels:
    p1afempy_b200.refinement.refine_nvb_device.)"""
    from .refinement import refine_nvb_device as run
    new_coords, out, new_boundaries, _ = run(coords, elements, marked, boundaries)
    return new_coords, out, new_boundaries
