"""CPU oracle for the P1 assembly hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy/scipy, the arithmetic of the reference
(leuraph/p1afempy v0.2.16) for the functions named in SURVEY.md section 8(a).
It exists so that the CUDA path can be checked on a GPU box where
``/root/reference`` is not present.  Nothing in ``p1afempy_b200`` may import
it; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do.

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` compares every
function below bit-for-bit with the real reference imported from
``/root/reference`` (through the import shims in ``oracle/shims``) whenever
that directory exists, and ``tests/test_oracle_golden.py`` compares it with
fixtures in ``tests/golden/`` that were produced by the real reference
(``tests/golden/make_golden.py``) including the MATLAB golden vectors the
reference's own tests hold (indicators.dat, mass_matrix_{I,J,D}.dat, area.dat,
the edge numbering constants of test_mesh.py).

The one piece of arithmetic that does NOT live in the reference is the
cubature table (third-party ``triangle-cubature``, unpinned in the reference's
pyproject.toml:20 and absent offline).  The rule is therefore *data* here:
every routine takes ``weights`` (n_qp,) and ``points`` (n_qp, 2) arrays; the
numeric values of DAYTAYLOR are "parity unpinned" (SURVEY.md section 8(c)).

Every function keeps the reference's floating point operation ORDER, because
the contract is bit-level where it can be (load vectors, estimator, local
matrices, off-diagonal CSR values) and 1e-12 where scipy's unstable per-row
sort makes the summation order undefined (CSR diagonals).
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix

# --------------------------------------------------------------------------
# geometry  (reference: p1afempy/mesh.py:10-51)
# --------------------------------------------------------------------------


def edge_vectors(coordinates, elements):
    """Vectors vertex0->vertex1 and vertex0->vertex2 per triangle.

    Follows mesh.py:47-49 (three row gathers, two subtractions).
    """
    base = coordinates[elements[:, 0], :]
    to_second = coordinates[elements[:, 1], :] - base
    to_third = coordinates[elements[:, 2], :] - base
    return to_second, to_third


def signed_area(coordinates, elements):
    """Signed triangle area, mesh.py:26-28: 0.5*(d21x*d31y - d21y*d31x)."""
    p, q = edge_vectors(coordinates, elements)
    return 0.5 * (p[:, 0] * q[:, 1] - p[:, 1] * q[:, 0])


# --------------------------------------------------------------------------
# triplet index streams
# --------------------------------------------------------------------------

def _triplet_indices_colmajor(elements):
    """(row, col) streams of solvers.py:34-37 / :171-174.

    Element-major stream of 9 entries; entry k of an element has
    row = vertex k%3, col = vertex k//3.
    """
    row_sel = [0, 1, 2, 0, 1, 2, 0, 1, 2]
    col_sel = [0, 0, 0, 1, 1, 1, 2, 2, 2]
    rows = elements[:, row_sel].reshape(-1)
    cols = elements[:, col_sel].reshape(-1)
    return rows, cols


def _triplet_indices_rowmajor(elements):
    """(row, col) streams of solvers.py:131-138 / :368-373.

    entry k of an element has row = vertex k//3, col = vertex k%3.
    """
    row_sel = [0, 0, 0, 1, 1, 1, 2, 2, 2]
    col_sel = [0, 1, 2, 0, 1, 2, 0, 1, 2]
    return elements[:, row_sel].reshape(-1), elements[:, col_sel].reshape(-1)


# --------------------------------------------------------------------------
# stiffness  (reference: p1afempy/solvers.py:15-47)
# --------------------------------------------------------------------------

def stiffness_triplets(coordinates, elements):
    """9*M triplets (I, J, V) of the P1 stiffness matrix, element-major.

    solvers.py:31      area4 = 4.*area
    solvers.py:41-43   a = (d21.d31)/area4, b = |d31|^2/area4, c = |d21|^2/area4
    solvers.py:45      local entries [(-2a+b)+c, a-b, a-c, a-b, b, -a, a-c, -a, c]
    """
    quad_area = 4. * signed_area(coordinates, elements)
    p, q = edge_vectors(coordinates, elements)
    a = (p[:, 0] * q[:, 0] + p[:, 1] * q[:, 1]) / quad_area
    b = (q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) / quad_area
    c = (p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]) / quad_area
    local = np.empty((elements.shape[0], 9))
    local[:, 0] = -2. * a + b + c
    local[:, 1] = a - b
    local[:, 2] = a - c
    local[:, 3] = a - b
    local[:, 4] = b
    local[:, 5] = -a
    local[:, 6] = a - c
    local[:, 7] = -a
    local[:, 8] = c
    rows, cols = _triplet_indices_colmajor(elements)
    return rows, cols, local.reshape(-1)


def stiffness_coo(coordinates, elements):
    """coo_matrix with INFERRED shape (solvers.py:46-47)."""
    i, j, v = stiffness_triplets(coordinates, elements)
    return coo_matrix((v, (i, j)))


# --------------------------------------------------------------------------
# mass  (reference: p1afempy/solvers.py:145-182)
# --------------------------------------------------------------------------

def mass_triplets(coordinates, elements):
    """(I, J, D) of solvers.py:156-182: area/6 on the local diagonal,
    area/12 elsewhere, same index streams as the stiffness matrix."""
    area = signed_area(coordinates, elements)
    on = area / 6.
    off = area / 12.
    local = np.empty((elements.shape[0], 9))
    for k in range(9):
        local[:, k] = on if k in (0, 4, 8) else off
    rows, cols = _triplet_indices_colmajor(elements)
    return rows, cols, local.reshape(-1)


def mass_coo(coordinates, elements):
    i, j, d = mass_triplets(coordinates, elements)
    return coo_matrix((d, (i, j)))


# --------------------------------------------------------------------------
# generalised stiffness  (reference: p1afempy/solvers.py:278-377)
# --------------------------------------------------------------------------

def quadrature_points(coordinates, elements, points):
    """x_q = z0 + eta*dz1 + xi*dz2 for every rule point (solvers.py:326).

    Returns an array (n_qp, M, 2)."""
    z0 = coordinates[elements[:, 0]]
    dz1 = coordinates[elements[:, 1]] - z0
    dz2 = coordinates[elements[:, 2]] - z0
    out = np.empty((len(points), elements.shape[0], 2))
    for q, (eta, xi) in enumerate(points):
        out[q] = z0 + eta * dz1 + xi * dz2
    return out


def general_stiffness_triplets(coordinates, elements, a11, a12, a21, a22,
                               weights, points):
    """Triplets of the generalised stiffness matrix.

    a11..a22 are callables (n,2)->(n,).  Operation order follows
    solvers.py:306-366 exactly (note the leading factor -1/(2*areas))."""
    z0 = coordinates[elements[:, 0]]
    dz1 = coordinates[elements[:, 1]] - z0
    dz2 = coordinates[elements[:, 2]] - z0
    alpha = dz2[:, 1]
    beta = -dz2[:, 0]
    gamma = -dz1[:, 1]
    delta = dz1[:, 0]
    areas = 0.5 * (dz1[:, 0] * dz2[:, 1] - dz1[:, 1] * dz2[:, 0])
    n_el = elements.shape[0]
    sa = np.zeros(n_el)
    sb = np.zeros(n_el)
    sc = np.zeros(n_el)
    sd = np.zeros(n_el)
    for w, (eta, xi) in zip(weights, points):
        xq = z0 + eta * dz1 + xi * dz2
        sa += w * a11(xq)
        sb += w * a12(xq)
        sc += w * a21(xq)
        sd += w * a22(xq)
    pre = -1. / (2. * areas)
    b11 = pre * (sa * alpha * alpha + sb * beta * alpha
                 + sc * alpha * beta + sd * beta * beta)
    b12 = pre * (sa * alpha * gamma + sb * delta * alpha
                 + sc * gamma * beta + sd * beta * delta)
    b21 = pre * (sa * alpha * gamma + sb * beta * gamma
                 + sc * alpha * delta + sd * beta * delta)
    b22 = pre * (sa * gamma * gamma + sb * delta * gamma
                 + sc * gamma * delta + sd * delta * delta)
    local = np.empty((n_el, 9))
    local[:, 0] = b11 + b12 + b21 + b22
    local[:, 1] = -(b11 + b21)
    local[:, 2] = -(b12 + b22)
    local[:, 3] = -(b11 + b12)
    local[:, 4] = b11
    local[:, 5] = b12
    local[:, 6] = -(b21 + b22)
    local[:, 7] = b21
    local[:, 8] = b22
    rows, cols = _triplet_indices_rowmajor(elements)
    return rows, cols, local.reshape(-1)


def general_stiffness_coo(coordinates, elements, a11, a12, a21, a22,
                          weights, points):
    """coo_matrix with EXPLICIT shape (N, N) (solvers.py:375-376)."""
    n = coordinates.shape[0]
    i, j, v = general_stiffness_triplets(coordinates, elements, a11, a12,
                                         a21, a22, weights, points)
    return coo_matrix((v, (i, j)), shape=(n, n))


# --------------------------------------------------------------------------
# weighted mass  (reference: p1afempy/solvers.py:50-142)
# --------------------------------------------------------------------------

def weighted_mass_triplets(coordinates, elements, current_iterate, phi,
                           weights, points):
    """Triplets of M_ij = int phi(u_h) phi_i phi_j.

    solvers.py:113-117  u_q = u1*h1 + u2*h2 + u3*h3
    solvers.py:121-129  a_kl += 2*areas*w*phi(u_q)*h_k*h_l (left to right)
    """
    z0 = coordinates[elements[:, 0]]
    dz1 = coordinates[elements[:, 1]] - z0
    dz2 = coordinates[elements[:, 2]] - z0
    areas = 0.5 * (dz1[:, 0] * dz2[:, 1] - dz1[:, 1] * dz2[:, 0])
    n_el = elements.shape[0]
    u1 = current_iterate[elements[:, 0]]
    u2 = current_iterate[elements[:, 1]]
    u3 = current_iterate[elements[:, 2]]
    local = np.zeros((9, n_el))
    for w, (eta, xi) in zip(weights, points):
        hat = (1 - eta - xi, eta, xi)
        uq = u1 * hat[0] + u2 * hat[1] + u3 * hat[2]
        pq = phi(uq)
        for k in range(3):
            for l in range(3):
                local[3 * k + l] += 2 * areas * w * pq * hat[k] * hat[l]
    rows, cols = _triplet_indices_rowmajor(elements)
    return rows, cols, np.ascontiguousarray(local.T).reshape(-1)


def weighted_mass_coo(coordinates, elements, current_iterate, phi,
                      weights, points):
    n = coordinates.shape[0]
    i, j, v = weighted_mass_triplets(coordinates, elements, current_iterate,
                                     phi, weights, points)
    return coo_matrix((v, (i, j)), shape=(n, n))


# --------------------------------------------------------------------------
# load vectors  (reference: p1afempy/solvers.py:380-426, 540-598)
# --------------------------------------------------------------------------

def load_vector_midpoint(coordinates, elements, f):
    """Legacy centroid rule, solvers.py:414-426.

    fsT = f(c0 + (d21+d31)/3); every vertex receives area4*fsT/12;
    accumulation order: all vertex-0 entries (by element), then vertex-1,
    then vertex-2 (bincount over elements.flatten('F'))."""
    quad_area = 4. * signed_area(coordinates, elements)
    p, q = edge_vectors(coordinates, elements)
    centroid = coordinates[elements[:, 0], :] + (p + q) / 3
    share = quad_area * f(centroid) / 12.
    idx = np.concatenate([elements[:, 0], elements[:, 1], elements[:, 2]])
    return np.bincount(idx, weights=np.concatenate([share, share, share]),
                       minlength=coordinates.shape[0])


def load_vector_cubature(coordinates, elements, f, weights, points):
    """solvers.py:540-598: L[m,j] += 2.*w*areas[m]*phi_j*f(x_q)[m], then one
    sequential bincount in (m, j) order."""
    areas = signed_area(coordinates, elements)
    n_el = elements.shape[0]
    z0 = coordinates[elements[:, 0], :]
    z1 = coordinates[elements[:, 1], :]
    z2 = coordinates[elements[:, 2], :]
    acc = np.zeros((n_el, 3))
    col_area = areas.reshape(n_el, 1)
    for w, (eta, xi) in zip(weights, points):
        hat = np.array([1. - eta - xi, eta, xi])
        xq = z0 + eta * (z1 - z0) + xi * (z2 - z0)
        fq = f(xq).reshape(n_el, 1)
        acc += 2. * w * col_area * hat * fq
    return np.bincount(elements.reshape(-1), weights=acc.reshape(-1),
                       minlength=coordinates.shape[0])


# --------------------------------------------------------------------------
# non-linear cousins  (reference: p1afempy/solvers.py:429-537) -- SURVEY.md
# section 8(f), rank 1 of the "next" rows
# --------------------------------------------------------------------------

def integrate_composition(f, u, coordinates, elements, weights, points):
    """int_Omega f(u_h) dx, solvers.py:457-472: per element
    L += w * f(u_q) * 2. * areas (left to right), then np.sum(L)."""
    areas = signed_area(coordinates, elements)
    acc = np.zeros(elements.shape[0], dtype=float)
    for w, (eta, xi) in zip(weights, points):
        uq = (u[elements[:, 0]] * (1 - eta - xi) + u[elements[:, 1]] * eta
              + u[elements[:, 2]] * xi)
        acc += w * f(uq) * 2. * areas
    return np.sum(acc)


def load_vector_composition(f, u, coordinates, elements, weights, points):
    """F_i = int f(u_h) phi_i, solvers.py:504-537: as the cubature load vector
    with f evaluated at u_q = u0*(1.-eta-xi) + u1*eta + u2*xi."""
    areas = signed_area(coordinates, elements)
    n_el = elements.shape[0]
    u0, u1, u2 = u[elements[:, 0]], u[elements[:, 1]], u[elements[:, 2]]
    acc = np.zeros((n_el, 3))
    col_area = areas.reshape(n_el, 1)
    for w, (eta, xi) in zip(weights, points):
        hat = np.array([1. - eta - xi, eta, xi])
        uq = u0 * (1. - eta - xi) + u1 * eta + u2 * xi
        acc += 2. * w * col_area * hat * f(uq).reshape(n_el, 1)
    return np.bincount(elements.reshape(-1), weights=acc.reshape(-1),
                       minlength=coordinates.shape[0])


# --------------------------------------------------------------------------
# edge numbering  (reference: p1afempy/mesh.py:87-161)
# --------------------------------------------------------------------------

def geometric_data(elements, boundaries):
    """element2edges (M,3), edge2nodes (E,2), [boundary_k -> edge ids].

    mesh.py:118-119  half-edge p = k*M + m goes e[m,k] -> e[m,(k+1)%3]
    mesh.py:124-130  every boundary is appended REVERSED
    mesh.py:133-136  edges are numbered in increasing p among entries I<J
    mesh.py:139-150  entries J<I inherit the number of their twin; the
                     reference matches them through two row-major sorted
                     COO->find passes, restated here with one lexsort each.
    """
    elements = np.asarray(elements)
    n_el = elements.shape[0]
    src = elements.reshape(-1, order='F').astype(np.int64)
    dst = elements[:, [1, 2, 0]].reshape(-1, order='F').astype(np.int64)
    starts = [3 * n_el]
    for bnd in boundaries:
        bnd = np.asarray(bnd)
        if bnd.size:
            src = np.concatenate([src, bnd[:, 1].astype(np.int64)])
            dst = np.concatenate([dst, bnd[:, 0].astype(np.int64)])
        starts.append(starts[-1] + bnd.shape[0])
    fwd = np.flatnonzero(src < dst)
    bwd = np.flatnonzero(dst < src)
    n_edges = fwd.size
    number = np.zeros(src.size, dtype=np.int64)
    number[fwd] = np.arange(n_edges)
    # twin matching: sort forward entries by (src,dst) and backward entries by
    # (dst,src); position-wise they are the same undirected edge.  (scipy's
    # find() returns row-major sorted entries and SUMS duplicates, so the
    # reference only works when each undirected edge appears exactly once in
    # each direction; the same requirement applies here.)
    order_f = np.lexsort((dst[fwd], src[fwd]))
    order_b = np.lexsort((src[bwd], dst[bwd]))
    if order_f.size != order_b.size:
        raise ValueError(
            "shape mismatch: every boundary edge must be listed in a boundary")
    number[bwd[order_b]] = number[fwd[order_f]]
    element2edges = number[:3 * n_el].reshape(n_el, 3, order='F')
    edge2nodes = np.column_stack((src[fwd], dst[fwd]))
    bnd2edges = [number[starts[k]:starts[k + 1]]
                 for k in range(len(boundaries))]
    return element2edges, edge2nodes, bnd2edges


# --------------------------------------------------------------------------
# residual estimator  (reference: p1afempy/indicators.py:7-86)
# --------------------------------------------------------------------------

def eta_r(x, coordinates, elements, dirichlet, neumann, f, g):
    """Squared residual indicators per element (indicators.py:42-86)."""
    element2edges, edge2nodes, bnd2edges = geometric_data(
        elements, [dirichlet, neumann])
    dbl_area = 2. * signed_area(coordinates, elements)
    p, q = edge_vectors(coordinates, elements)
    du1 = x[elements[:, 1]] - x[elements[:, 0]]
    du2 = x[elements[:, 2]] - x[elements[:, 0]]
    curl_x = (q[:, 0] * du1 - p[:, 0] * du2) / dbl_area
    curl_y = (q[:, 1] * du1 - p[:, 1] * du2) / dbl_area
    dn21 = p[:, 0] * curl_x + p[:, 1] * curl_y
    dn13 = -(q[:, 0] * curl_x + q[:, 1] * curl_y)
    dn32 = -(dn13 + dn21)
    jump = np.bincount(element2edges.reshape(-1, order='F'),
                       np.concatenate([dn21, dn32, dn13]),
                       minlength=edge2nodes.shape[0])
    neumann = np.asarray(neumann)
    if neumann.size > 0:
        n1 = coordinates[neumann[:, 0], :]
        n2 = coordinates[neumann[:, 1], :]
        gm = g((n1 + n2) / 2.)
        ne = bnd2edges[1]
        dx = n2 - n1
        length = np.sqrt(np.square(dx[:, 0]) + np.square(dx[:, 1]))
        jump[ne] = jump[ne] - length * gm
    jump[bnd2edges[0]] = 0
    sq = np.square(jump[element2edges])
    eta = sq[:, 0] + sq[:, 1] + sq[:, 2]
    centroid = coordinates[elements[:, 0]] + (p + q) / 3.
    return eta + np.square(0.5 * dbl_area * f(centroid))


# --------------------------------------------------------------------------
# CSR canonicalisation = what "to CSR" means (scipy/_coo.py:368-439,
# scipy/_compressed.py:1072-1130).  scipy itself is part of the oracle.
# --------------------------------------------------------------------------

def to_csr(coo):
    return coo.tocsr()
