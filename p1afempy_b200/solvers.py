"""Drop-in for the assembly builders of ``p1afempy.solvers`` -- same names,
parameter names, order and defaults (/root/reference/p1afempy/solvers.py:15-598).

numpy arrays in; ``scipy.sparse.csr_matrix`` (== ``reference(...).tocsr()``)
or ``np.ndarray`` out.  Everything between is CUDA (libp1b200.so); user
callbacks (``f``, ``a_ij``, ``phi``) are evaluated on the host on exactly the
points the reference would pass them, everything else stays on the device.
"""
from __future__ import annotations

import numpy as np

import torch

from . import devarray, options
from .cubature import resolve_rule
from .device import DeviceMesh, to_host

__all__ = ['get_stiffness_matrix', 'get_mass_matrix', 'get_mass_matrix_elements',
           'get_general_stiffness_matrix', 'get_weighted_mass_matrix',
           'get_right_hand_side', 'get_right_hand_side_using_quadrature_rule',
           'integrate_composition_nonlinear_with_fem',
           'get_load_vector_of_composition_nonlinear_with_fem',
           'apply_neumann', 'solve_laplace']


def _values_on_host(fn, points_host, n):
    """Calls a user callback like the reference does (one (n, d) array in,
    (n,) out) and normalises what it returns (scalar, list, (n,1) ...)."""
    out = np.asarray(fn(points_host), dtype=np.float64)
    if out.ndim == 0:
        out = np.full(n, float(out))
    out = out.reshape(-1)
    if out.shape[0] != n:
        raise ValueError('callback returned %d values for %d points'
                         % (out.shape[0], n))
    return out


def _callback_at_points(fns, xq):
    """fns evaluated at every quadrature point set; xq: CUDA (n_qp, M, 2) or
    (n_qp, M).  -> list of (n_qp, M) arrays, one per callback: CUDA tensors in
    'device' mode (options.callbacks), host arrays otherwise."""
    n_qp, m = xq.shape[0], xq.shape[1]
    outs = [None] * len(fns)
    todo = list(range(len(fns)))
    if options.callbacks == 'device':
        for j in list(todo):
            vals = []
            for q in range(n_qp):
                v = devarray.evaluate(fns[j], xq[q], m)
                if v is None:
                    break
                vals.append(v)
            else:
                outs[j] = torch.stack(vals) if vals else xq.new_empty((0, m))
                todo.remove(j)
    if todo:   # host evaluation of the remaining callbacks (only the callbacks)
        pts = to_host(xq)
        for j in todo:
            outs[j] = np.empty((n_qp, m))
            for q in range(n_qp):
                outs[j][q] = _values_on_host(fns[j], pts[q], m)
    return outs


def _callback_at(fn, points):
    """One callback on one CUDA point/value array (n, 2) or (n,) -> (n,)."""
    return _callback_at_points((fn,), points[None])[0][0]


def _need_inferable(mesh):
    if mesh.n_elements == 0:
        raise ValueError('cannot infer dimensions from zero sized index arrays')


def get_stiffness_matrix(coordinates, elements):
    """P1 stiffness matrix (solvers.py:15-47) as CSR."""
    with DeviceMesh(coordinates, elements) as mesh:
        _need_inferable(mesh)
        return mesh.to_scipy(*mesh.stiffness_csr(), inferred_shape=True)


def get_mass_matrix(coordinates, elements):
    """P1 mass matrix (solvers.py:145-153) as CSR."""
    with DeviceMesh(coordinates, elements) as mesh:
        _need_inferable(mesh)
        return mesh.to_scipy(*mesh.mass_csr(), inferred_shape=True)


def get_mass_matrix_elements(coordinates, elements):
    """(I, J, D) triplets in the reference's order (solvers.py:156-182)."""
    with DeviceMesh(coordinates, elements) as mesh:
        rows, cols, vals = mesh.triplets(stiffness=False)
        kind = np.asarray(elements).dtype
        i, j = to_host(rows), to_host(cols)
        if kind.kind in 'iu' and kind != np.int64:
            i, j = i.astype(kind), j.astype(kind)
        return i, j, to_host(vals)


def get_general_stiffness_matrix(coordinates, elements, a_11, a_12, a_21, a_22,
                                 cubature_rule):
    """Stiffness matrix of -div(A(x) grad u) (solvers.py:278-377) as CSR with
    explicit shape (N, N)."""
    weights, points = resolve_rule(cubature_rule)
    with DeviceMesh(coordinates, elements) as mesh:
        xq = mesh.quadrature_points(points)
        q11, q12, q21, q22 = _callback_at_points((a_11, a_12, a_21, a_22), xq)
        del xq
        return mesh.to_scipy(*mesh.general_stiffness_csr(q11, q12, q21, q22, weights),
                             inferred_shape=False)


def get_weighted_mass_matrix(coordinates, elements, current_iterate, phi,
                             cubature_rule):
    """M_ij = int phi(u_h) phi_i phi_j (solvers.py:50-142) as CSR (N, N)."""
    weights, points = resolve_rule(cubature_rule)
    with DeviceMesh(coordinates, elements) as mesh:
        u = np.asarray(current_iterate, dtype=np.float64).reshape(-1)
        top = int(mesh.elements.max().item()) if mesh.n_elements else -1
        if u.shape[0] < top + 1:
            raise IndexError('current_iterate has %d entries, elements reference '
                             'node %d' % (u.shape[0], top))
        if u.shape[0] != mesh.n_nodes:   # numpy would only read what it needs
            u = np.concatenate([u, np.zeros(max(0, mesh.n_nodes - u.shape[0]))])[:mesh.n_nodes]
        (phiq,) = _callback_at_points((phi,), mesh.nodal_at_quadrature(u, points))
        return mesh.to_scipy(*mesh.weighted_mass_csr(phiq, weights, points),
                             inferred_shape=False)


def get_right_hand_side(coordinates, elements, f, cubature_rule=None):
    """Load vector (solvers.py:380-426); legacy centroid rule when
    ``cubature_rule`` is None."""
    if cubature_rule is not None:
        return get_right_hand_side_using_quadrature_rule(
            coordinates=coordinates, elements=elements, f=f,
            cubature_rule=cubature_rule)
    with DeviceMesh(coordinates, elements) as mesh:
        fc = _callback_at(f, mesh.centroids())
        return to_host(mesh.load_vector_midpoint(fc))


def get_right_hand_side_using_quadrature_rule(coordinates, elements, f,
                                              cubature_rule):
    """Load vector with a cubature rule (solvers.py:540-598)."""
    weights, points = resolve_rule(cubature_rule)
    with DeviceMesh(coordinates, elements) as mesh:
        (fq,) = _callback_at_points((f,), mesh.quadrature_points(points))
        return to_host(mesh.load_vector(fq, weights, points))


def _nodal_checked(mesh, u):
    u = np.asarray(u, dtype=np.float64).reshape(-1)
    top = int(mesh.elements.max().item()) if mesh.n_elements else -1
    if u.shape[0] < top + 1:
        raise IndexError('u has %d entries, elements reference node %d' % (u.shape[0], top))
    if u.shape[0] != mesh.n_nodes:
        u = np.concatenate([u, np.zeros(max(0, mesh.n_nodes - u.shape[0]))])[:mesh.n_nodes]
    return u


def integrate_composition_nonlinear_with_fem(f, u, coordinates, elements, cubature_rule):
    """int_Omega f(u_h) dx for a P1 function u_h (solvers.py:429-472)."""
    weights, points = resolve_rule(cubature_rule)
    with DeviceMesh(coordinates, elements) as mesh:
        uq = mesh.nodal_at_quadrature(_nodal_checked(mesh, u), points)
        (fq,) = _callback_at_points((f,), uq)
        return mesh.integrate(fq, weights)


def get_load_vector_of_composition_nonlinear_with_fem(f, u, coordinates, elements,
                                                      cubature_rule):
    """F_i = int_Omega f(u_h) phi_i dx (solvers.py:475-537)."""
    weights, points = resolve_rule(cubature_rule)
    with DeviceMesh(coordinates, elements) as mesh:
        uq = mesh.nodal_at_quadrature(_nodal_checked(mesh, u), points)
        (fq,) = _callback_at_points((f,), uq)
        return to_host(mesh.load_vector(fq, weights, points))


# ---------------------------------------------------------------------------
# The reference's driver ("next" rank 2: the consumer side of
# the path, /root/reference/p1afempy/solvers.py:601-697).  Assembly, load vector,
# Neumann load, Dirichlet lift and energy run on the device on the device-resident
# CSR; ``solver='direct'`` (default) then hands the free-node system to scipy's
# direct solver like the reference does, ``solver='pcg'`` solves it on the GPU
# (Jacobi-preconditioned CG, p1_pcg), so that the matrix never crosses PCIe.
# ---------------------------------------------------------------------------

def _neumann_on_device(mesh, neumann_t, g, b_t):
    """b + Neumann contributions on the device (p1_boundary_edges, p1_apply_neumann)."""
    import ctypes as C                                   # noqa: F401
    from . import _lib
    from .device import _ptr, _stream
    _, mid = mesh.boundary_edges(neumann_t)
    gm = _callback_at(g, mid)
    if not isinstance(gm, torch.Tensor):
        gm = torch.from_numpy(np.ascontiguousarray(gm, dtype=np.float64)).to(b_t.device)
    out = torch.empty_like(b_t)
    _lib.check(mesh._lib.p1_apply_neumann(
        mesh.ctx, _ptr(mesh.coords), mesh.n_nodes, _ptr(neumann_t), int(neumann_t.shape[0]),
        _ptr(gm.contiguous()), _ptr(b_t), _ptr(out), _stream(mesh.device)))
    return out


def apply_neumann(neumann_bc, coordinates, g, b):
    """b + Neumann contributions (solvers.py:601-618): every boundary edge adds
    |E| g(midpoint) / 2 to both of its nodes, summed like the reference's bincount
    (first endpoints, then second endpoints)."""
    from .device import to_device_f64, to_device_indices
    neumann_bc = np.asarray(neumann_bc).reshape(-1, 2)
    b = np.asarray(b, dtype=np.float64)
    with DeviceMesh(coordinates, np.zeros((0, 3), dtype=np.int64), check=False) as mesh:
        if mesh.n_nodes < b.size:
            raise ValueError('apply_neumann: b has more entries than there are coordinates')
        n_t = to_device_indices(neumann_bc, mesh.device, mesh.ctx).reshape(-1, 2)
        b_t = to_device_f64(np.concatenate([b, np.zeros(mesh.n_nodes - b.size)]), mesh.device)
        return to_host(_neumann_on_device(mesh, n_t, g, b_t))[:b.size]


def solve_laplace(coordinates, elements, dirichlet, neumann, f, g, uD,
                  cubature_rule=None, solver='direct', tol=1e-12, max_iter=None,
                  return_info=False):
    """-Laplace u = f with mixed boundary conditions (solvers.py:621-697).
    Returns (x, energy) [, info with solver='pcg' and return_info].

    solver='direct': scipy's sparse direct solver on the free nodes, as the reference.
    solver='pcg'   : conjugate gradient on the device; the matrix stays on the GPU
                     (tol: relative residual, max_iter: default 10 N)."""
    import ctypes as C
    from . import _lib
    from .device import _ptr, _stream, to_device_f64, to_device_indices
    if solver not in ('direct', 'pcg'):
        raise ValueError("solver must be 'direct' or 'pcg'")
    coordinates = np.asarray(coordinates, dtype=np.float64)
    n = coordinates.shape[0]
    weights = points = None
    if cubature_rule is not None:
        weights, points = resolve_rule(cubature_rule)
    with DeviceMesh(coordinates, elements) as mesh:
        _need_inferable(mesh)
        dev, stream = torch.device('cuda', mesh.device), _stream(mesh.device)
        ip, ix, data = mesh.stiffness_csr()
        if mesh.max_index + 1 < n:             # trailing unused nodes (inferred shape)
            raise ValueError('solve_laplace: every node must belong to an element')
        # prescribed values at the Dirichlet nodes
        fixed = np.unique(np.asarray(dirichlet))
        fixed_t = torch.from_numpy(fixed.astype(np.int64)).to(dev)
        ud = _callback_at(uD, mesh.coords[fixed_t])
        if not isinstance(ud, torch.Tensor):
            ud = torch.from_numpy(np.ascontiguousarray(ud, dtype=np.float64)).to(dev)
        x = torch.zeros(n, dtype=torch.float64, device=dev)
        x[fixed_t] = ud
        # right-hand side: load vector - A x_D (+ Neumann)
        if points is None:
            b = mesh.load_vector_midpoint(_callback_at(f, mesh.centroids()))
        else:
            (fq,) = _callback_at_points((f,), mesh.quadrature_points(points))
            b = mesh.load_vector(fq, weights, points)
        if solver == 'direct':      # (p1_pcg lifts x_D itself: its first residual is b - A x)
            _lib.check(mesh._lib.p1_csr_spmv(mesh.ctx, n, _ptr(ip), _ptr(ix), _ptr(data),
                                             _ptr(x), -1.0, _ptr(b), _ptr(b), stream))
        neumann = np.asarray(neumann)
        if neumann.size > 0:
            n_t = to_device_indices(neumann.reshape(-1, 2), mesh.device, mesh.ctx).reshape(-1, 2)
            b = _neumann_on_device(mesh, n_t, g, b)
        if solver == 'pcg':
            mask = torch.zeros(n, dtype=torch.uint8, device=dev)
            mask[fixed_t] = 1
            its, res = C.c_int(), C.c_double()
            _lib.check(mesh._lib.p1_pcg(mesh.ctx, n, _ptr(ip), _ptr(ix), _ptr(data), _ptr(b),
                                        _ptr(mask), _ptr(x), float(tol),
                                        int(max_iter if max_iter is not None else 10 * n),
                                        C.byref(its), C.byref(res), stream))
            ax = torch.empty_like(x)
            _lib.check(mesh._lib.p1_csr_spmv(mesh.ctx, n, _ptr(ip), _ptr(ix), _ptr(data),
                                             _ptr(x), 1.0, None, _ptr(ax), stream))
            energy = C.c_double()
            _lib.check(mesh._lib.p1_dot(mesh.ctx, n, _ptr(x), _ptr(ax), C.byref(energy), stream))
            out = (to_host(x), energy.value)
            return out + ({'iterations': its.value, 'relative_residual': res.value},) \
                if return_info else out
        # the reference's direct solve on the host (the only step that needs A there)
        from scipy.sparse import csc_matrix
        from scipy.sparse.linalg import spsolve
        a = csc_matrix(mesh.to_scipy(ip, ix, data, inferred_shape=True))
        xh, bh = to_host(x), to_host(b)
    free = np.setdiff1d(np.arange(0, n), fixed, assume_unique=True)
    xh[free] = spsolve(a[free, :][:, free], bh[free], use_umfpack=True)
    energy = xh.dot(a.dot(xh))
    return xh, energy
