"""Parity of the CUDA path (through the drop-in functions, i.e. through the
C ABI) with the oracle and with the committed golden fixtures.

Contract: CSR index arrays bit-exact (dtype and shape
included); values within 1e-12 of the row maximum; load vectors, estimator,
edge numbering and triplets bit-exact.
"""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from oracle import p1_oracle as orc
from p1afempy_b200 import cubature, meshgen
import helpers as H
import helpers_laplace as L

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.fixture(scope='module')
def api():
    from p1afempy_b200 import solvers, mesh, indicators
    return solvers, mesh, indicators


def load(name):
    return np.load(os.path.join(H.GOLDEN, name + '.npz'))


def csr_of(g, prefix):
    return csr_matrix((g[prefix + '_data'], g[prefix + '_indices'],
                       g[prefix + '_indptr']), shape=tuple(g[prefix + '_shape']))


MESHES = {
    'perturbed_s5': lambda: H.perturbed_square(5),
    'graded': lambda: H.graded_square(),
    'lshape_u4': lambda: meshgen.refine_uniform(*meshgen.l_shape(), times=4),
    'two_triangles': lambda: meshgen.two_triangle_square(),
    'uniform_s7': lambda: meshgen.unit_square(7),
}


@pytest.fixture(scope='module', params=list(MESHES))
def mesh_case(request):
    return MESHES[request.param]()


def test_stiffness_vs_oracle(api, mesh_case):
    c, e, _ = mesh_case
    a = api[0].get_stiffness_matrix(c, e)
    H.csr_values_close(a, orc.stiffness_coo(c, e).tocsr(), RTOL)
    assert a.has_sorted_indices


def test_mass_vs_oracle(api, mesh_case):
    c, e, _ = mesh_case
    H.csr_values_close(api[0].get_mass_matrix(c, e), orc.mass_coo(c, e).tocsr(), RTOL)
    i, j, d = api[0].get_mass_matrix_elements(c, e)
    oi, oj, od = orc.mass_triplets(c, e)
    assert np.array_equal(i, oi) and np.array_equal(j, oj)
    assert np.array_equal(d, od)                      # bit-exact


@pytest.mark.parametrize('rule', ['MIDPOINT', 'DAYTAYLOR'])
def test_general_stiffness_vs_oracle(api, mesh_case, rule):
    c, e, _ = mesh_case
    r = cubature.CubatureRuleEnum[rule]
    w, p = cubature.resolve_rule(r)
    a = api[0].get_general_stiffness_matrix(c, e, H.a11, H.a12, H.a21, H.a22, r)
    ref = orc.general_stiffness_coo(c, e, H.a11, H.a12, H.a21, H.a22, w, p).tocsr()
    H.csr_values_close(a, ref, RTOL)


@pytest.mark.parametrize('rule', ['MIDPOINT', 'DAYTAYLOR'])
def test_weighted_mass_vs_oracle(api, mesh_case, rule):
    c, e, _ = mesh_case
    r = cubature.CubatureRuleEnum[rule]
    w, p = cubature.resolve_rule(r)
    u = H.u_nodal(c)
    a = api[0].get_weighted_mass_matrix(c, e, u, H.phi, r)
    H.csr_values_close(a, orc.weighted_mass_coo(c, e, u, H.phi, w, p).tocsr(), RTOL)


def test_load_vectors_bit_exact(api, mesh_case):
    c, e, _ = mesh_case
    assert np.array_equal(api[0].get_right_hand_side(c, e, H.f_load),
                          orc.load_vector_midpoint(c, e, H.f_load))
    for rule in ('MIDPOINT', 'SMPLX1', 'DAYTAYLOR'):
        r = cubature.CubatureRuleEnum[rule]
        w, p = cubature.resolve_rule(r)
        mine = api[0].get_right_hand_side(c, e, H.f_load, cubature_rule=r)
        assert np.array_equal(mine, orc.load_vector_cubature(c, e, H.f_load, w, p))
        # the rule is data: a (weights, points) pair works as well
        assert np.array_equal(
            api[0].get_right_hand_side_using_quadrature_rule(c, e, H.f_load, (w, p)),
            mine)


def test_geometry_bit_exact(api, mesh_case):
    c, e, _ = mesh_case
    assert np.array_equal(api[1].get_area(c, e), orc.signed_area(c, e))
    d21, d31 = api[1].get_directional_vectors(c, e)
    o21, o31 = orc.edge_vectors(c, e)
    assert np.array_equal(d21, o21) and np.array_equal(d31, o31)


def _split(b):
    allb = np.vstack(b)
    cut = (3 * allb.shape[0]) // 5
    return allb[:cut], allb[cut:]


def test_geometric_data_bit_exact(api, mesh_case):
    _, e, b = mesh_case
    parts = list(_split(b))
    e2e, e2n, b2e = api[1].provide_geometric_data(e, parts)
    o2e, o2n, ob2e = orc.geometric_data(e, parts)
    assert e2e.dtype == o2e.dtype == np.int64
    assert np.array_equal(e2e, o2e) and np.array_equal(e2n, o2n)
    assert len(b2e) == 2
    for x, y in zip(b2e, ob2e):
        assert np.array_equal(x, y)


@pytest.mark.parametrize('with_neumann', [True, False])
def test_eta_r_bit_exact(api, mesh_case, with_neumann):
    c, e, b = mesh_case
    if with_neumann:
        dirichlet, neumann = _split(b)
    else:
        dirichlet, neumann = np.vstack(b), np.zeros((0, 2), dtype=int)
    x = H.u_nodal(c) + 0.1 * c[:, 0]
    mine = api[2].compute_eta_r(x, c, e, dirichlet, neumann, H.f_load, H.g_neu)
    ref = orc.eta_r(x, c, e, dirichlet, neumann, H.f_load, H.g_neu)
    assert np.array_equal(mine, ref)


# ------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize('case', ['perturbed_s4', 'graded', 'lshape_u3'])
def test_against_reference_fixtures(api, case):
    """Outputs of the REAL reference, committed under tests/golden."""
    solvers, mesh, indicators = api
    g = load(case)
    c, e, u, x = g['coordinates'], g['elements'], g['u'], g['x']
    H.csr_values_close(solvers.get_stiffness_matrix(c, e), csr_of(g, 'stiff'), RTOL)
    H.csr_values_close(solvers.get_mass_matrix(c, e), csr_of(g, 'mass'), RTOL)
    i, j, d = solvers.get_mass_matrix_elements(c, e)
    assert np.array_equal(i, g['mass_I']) and np.array_equal(j, g['mass_J'])
    assert np.array_equal(d, g['mass_D'])
    for r in ('MIDPOINT', 'DAYTAYLOR'):
        rule = (g['rule_%s_w' % r], g['rule_%s_p' % r])
        H.csr_values_close(
            solvers.get_general_stiffness_matrix(c, e, H.a11, H.a12, H.a21, H.a22, rule),
            csr_of(g, 'gen_' + r), RTOL)
        H.csr_values_close(solvers.get_weighted_mass_matrix(c, e, u, H.phi, rule),
                           csr_of(g, 'wmass_' + r), RTOL)
        assert np.array_equal(solvers.get_right_hand_side(c, e, H.f_load, rule),
                              g['rhs_' + r])
    assert np.array_equal(solvers.get_right_hand_side(c, e, H.f_load), g['rhs_legacy'])
    e2e, e2n, b2e = mesh.provide_geometric_data(e, [g['dirichlet'], g['neumann']])
    assert np.array_equal(e2e, g['element2edges'])
    assert np.array_equal(e2n, g['edge2nodes'])
    assert np.array_equal(b2e[0], g['dir2edges'])
    assert np.array_equal(b2e[1], g['neu2edges'])
    assert np.array_equal(
        indicators.compute_eta_r(x, c, e, g['dirichlet'], g['neumann'], H.f_load, H.g_neu),
        g['eta'])
    assert np.array_equal(mesh.get_area(c, e), g['area'])


def test_laplace_example_matlab(api):
    """test_indicators.py:12-55 / test_solver.py:20-57 of the reference."""
    solvers, _, indicators = api
    g = load('laplace_example')
    eta = indicators.compute_eta_r(g['x_reference'], g['coordinates'], g['elements'],
                                   g['dirichlet'], g['neumann'], L.f, L.g)
    assert np.array_equal(eta, g['eta_reference'])
    assert np.allclose(eta, g['eta_matlab'])
    a = solvers.get_stiffness_matrix(g['coordinates3'], g['elements3'])
    xm = g['x_matlab']
    assert np.isclose(xm.dot(a.dot(xm)), g['energy_matlab'])


def test_ahw_mass_triplets_matlab(api):
    """test_solver.py:59-86: triplet order pinned by MATLAB data."""
    g = load('ahw_example')
    i, j, d = api[0].get_mass_matrix_elements(g['coordinates'], g['elements'])
    assert np.array_equal(i, g['mass_I_matlab']) and np.array_equal(j, g['mass_J_matlab'])
    assert np.allclose(d, g['mass_D_matlab'])
    assert np.allclose(api[1].get_area(g['coordinates'], g['elements']), g['area_matlab'])


def test_edge_numbering_known_answers(api):
    """test_mesh.py:29-112 (uint32 inputs, as io_helpers returns them)."""
    g = load('edge_numbering')
    for pre, nb in (('sq', 2), ('l', 3)):
        bnds = [g['%s_b%d' % (pre, k)] for k in range(nb)]
        e2e, e2n, b2e = api[1].provide_geometric_data(g[pre + '_elements'], bnds)
        assert np.array_equal(e2e, g[pre + '_element2edges'])
        assert np.array_equal(e2n, g[pre + '_edge2nodes'])
        for k in range(nb):
            assert np.array_equal(b2e[k], g['%s_b2e%d' % (pre, k)])


def test_sanity_energy_exactly_four(api):
    """test_sanity_check.py:80-119: assertEqual(4.0, x.A.x)."""
    g = load('sanity_check')
    for level in range(4):
        c, e = g['c%d' % level], g['e%d' % level]
        x = L.sanity_u(c)
        a = api[0].get_stiffness_matrix(c, e)
        assert x.dot(a.dot(x)) == 4.0


def test_general_minus_identity_equals_stiffness(api):
    """test_general_stiffness_matrices.py:22-59; callbacks build their result
    from .shape[0] on the host."""
    c, e, _ = H.perturbed_square(4)
    minus_one = lambda r: -np.ones(r.shape[0])   # noqa: E731
    zero = lambda r: np.zeros(r.shape[0])        # noqa: E731
    gen = api[0].get_general_stiffness_matrix(
        c, e, minus_one, zero, zero, minus_one, cubature.CubatureRuleEnum.DAYTAYLOR)
    assert np.allclose(gen.toarray(), api[0].get_stiffness_matrix(c, e).toarray())


# ---------------------------------------------------------------- edge cases
def test_index_dtypes_and_unused_nodes(api):
    c, e, _ = H.perturbed_square(3)
    ref = orc.stiffness_coo(c, e).tocsr()
    for dt in (np.int32, np.int64, np.uint32, np.int16):
        H.csr_values_close(api[0].get_stiffness_matrix(c, e.astype(dt)), ref, RTOL)
    # trailing unused nodes: inferred shape shrinks, explicit shape does not
    c2 = np.vstack([c, [[5., 5.], [6., 6.]]])
    a = api[0].get_stiffness_matrix(c2, e)
    assert a.shape == ref.shape
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.MIDPOINT)
    gen = api[0].get_general_stiffness_matrix(c2, e, H.a11, H.a12, H.a21, H.a22, (w, p))
    gref = orc.general_stiffness_coo(c2, e, H.a11, H.a12, H.a21, H.a22, w, p).tocsr()
    H.csr_values_close(gen, gref, RTOL)
    assert gen.shape == (c2.shape[0],) * 2
    b = api[0].get_right_hand_side(c2, e, H.f_load)
    assert np.array_equal(b, orc.load_vector_midpoint(c2, e, H.f_load))
    # an unused node in the middle -> empty row
    e3 = e.copy()
    e3[e3 >= 7] += 1
    c3 = np.insert(c, 7, [[9., 9.]], axis=0)
    H.csr_values_close(api[0].get_stiffness_matrix(c3, e3),
                       orc.stiffness_coo(c3, e3).tocsr(), RTOL)


def test_out_of_range_index_raises(api):
    c, e, _ = meshgen.unit_square(2)
    bad = e.copy()
    bad[3, 1] = c.shape[0]
    with pytest.raises(IndexError):
        api[0].get_stiffness_matrix(c, bad)
    bad[3, 1] = -1
    with pytest.raises(IndexError):
        api[0].get_right_hand_side(c, bad, H.f_load)


def test_incomplete_boundary_raises(api):
    c, e, b = meshgen.unit_square(3)
    partial = [np.vstack(b)[:-3]]
    with pytest.raises(ValueError):
        api[1].provide_geometric_data(e, partial)
    with pytest.raises(ValueError):
        api[2].compute_eta_r(H.u_nodal(c), c, e, partial[0],
                             np.zeros((0, 2), dtype=int), H.f_load, H.g_neu)


def test_high_valence_fan(api):
    """One node shared by 40 triangles (more than the fast path's DEG_MAX):
    exercises the block-per-row path."""
    n = 40
    ang = 2 * np.pi * np.arange(n) / n
    c = np.vstack([[0., 0.], np.column_stack([np.cos(ang), np.sin(ang)])])
    c[1:] *= (1 + 0.1 * np.sin(5 * ang))[:, None]
    e = np.column_stack([np.zeros(n, dtype=int), 1 + np.arange(n),
                         1 + (np.arange(n) + 1) % n])
    H.csr_values_close(api[0].get_stiffness_matrix(c, e),
                       orc.stiffness_coo(c, e).tocsr(), RTOL)
    H.csr_values_close(api[0].get_mass_matrix(c, e), orc.mass_coo(c, e).tocsr(), RTOL)
    assert np.array_equal(api[0].get_right_hand_side(c, e, H.f_load),
                          orc.load_vector_midpoint(c, e, H.f_load))
    rim = np.column_stack([1 + np.arange(n), 1 + (np.arange(n) + 1) % n])
    e2e, e2n, b2e = api[1].provide_geometric_data(e, [rim])
    o2e, o2n, ob2e = orc.geometric_data(e, [rim])
    assert np.array_equal(e2e, o2e) and np.array_equal(e2n, o2n)
    assert np.array_equal(b2e[0], ob2e[0])


def test_non_manifold_still_assembles(api):
    """Three triangles on one edge: not a mesh, but the reference just sums
    triplets; so must the assembly."""
    c = np.array([[0., 0.], [1., 0.], [0., 1.], [1., 1.], [.5, -1.]])
    e = np.array([[0, 1, 2], [1, 0, 4], [0, 1, 3]])
    H.csr_values_close(api[0].get_stiffness_matrix(c, e),
                       orc.stiffness_coo(c, e).tocsr(), RTOL)


def test_large_uniform_invariants(api):
    """S(9) (1M triangles, config C1): full parity against the oracle plus
    size-independent properties (row sums of the stiffness matrix vanish,
    symmetry, total mass = area of the domain)."""
    c, e, _ = meshgen.unit_square(9)
    a = api[0].get_stiffness_matrix(c, e)
    H.csr_values_close(a, orc.stiffness_coo(c, e).tocsr(), RTOL)
    assert np.abs(a @ np.ones(a.shape[0])).max() < 1e-9
    assert abs(a - a.T).max() == 0.0
    m = api[0].get_mass_matrix(c, e)
    assert abs(m.sum() - 1.0) < 1e-12


def test_pinched_vertex_falls_back_to_exact_plan(api):
    """Two closed fans that share only one vertex: the default plan's
    closed-fan shortcut is contradicted while filling (P1_ERETRY) and the
    wrapper re-plans exactly.  A degenerate element (repeated vertex) too."""
    def fan(center, first, n, c0):
        ang = 2 * np.pi * np.arange(n) / n
        pts = np.column_stack([c0[0] + np.cos(ang), c0[1] + np.sin(ang)])
        ids = first + np.arange(n)
        return pts, np.column_stack([np.full(n, center), ids, np.roll(ids, -1)])
    p1, e1 = fan(0, 1, 6, (-1., 0.))
    p2, e2 = fan(0, 7, 6, (1., 0.))
    c = np.vstack([[0., 0.], p1 * 0.9, p2 * 0.9])
    c[1:7] += [-0.2, 0.]
    c[7:] += [0.2, 0.]
    e = np.vstack([e1, e2])
    H.csr_values_close(api[0].get_stiffness_matrix(c, e),
                       orc.stiffness_coo(c, e).tocsr(), RTOL)
    H.csr_values_close(api[0].get_mass_matrix(c, e), orc.mass_coo(c, e).tocsr(), RTOL)
    c3, e3, _ = meshgen.unit_square(2)
    e3 = np.vstack([e3, [[3, 3, 7]]])           # zero-area element with a repeated node
    with np.errstate(all='ignore'):
        ref = orc.mass_coo(c3, e3).tocsr()
        H.csr_values_close(api[0].get_mass_matrix(c3, e3), ref, RTOL)


def test_default_callbacks_run_on_the_device(api):
    """The package default evaluates callbacks on the device; arithmetic callbacks
    (no transcendental) then still give the host mode's bits."""
    import importlib
    from p1afempy_b200 import options
    assert importlib.reload(importlib.import_module('p1afempy_b200.options')).callbacks == \
        os.environ.get('P1_CALLBACKS', 'device')
    solvers = api[0]
    c, e, _ = H.perturbed_square(4)
    rule = cubature.CubatureRuleEnum.DAYTAYLOR
    poly = lambda r: 2. + r[:, 0] * r[:, 1] - 0.5 * r[:, 1] ** 2 + r[:, 0] / (3. + r[:, 1])   # noqa: E731
    out = []
    for mode in ('host', 'device'):
        options.callbacks = mode
        out.append((solvers.get_general_stiffness_matrix(c, e, H.a11, H.a12, H.a21, poly, rule),
                    solvers.get_weighted_mass_matrix(c, e, H.u_nodal(c), H.phi, rule),
                    solvers.get_right_hand_side(c, e, poly, cubature_rule=rule),
                    solvers.get_right_hand_side(c, e, poly)))
    options.callbacks = 'host'
    for x, y in zip(*out):
        if hasattr(x, 'data') and hasattr(x, 'indptr'):
            assert np.array_equal(x.indices, y.indices) and np.array_equal(x.data, y.data)
        else:
            assert np.array_equal(x, y)


def test_device_callbacks(api):
    """options.callbacks='device': the reference's callbacks run unchanged on a
    device-resident ndarray stand-in; values agree to 1e-12 (CUDA libm instead
    of glibc), index arrays stay bit-exact; callbacks the stand-in cannot
    express (boolean-mask assignment: helpers_laplace.g) fall back to the host."""
    from p1afempy_b200 import options
    solvers, mesh, indicators = api
    c, e, b = H.perturbed_square(5)
    rule = cubature.CubatureRuleEnum.DAYTAYLOR
    w, p = cubature.resolve_rule(rule)
    u = H.u_nodal(c)
    options.callbacks = 'device'
    try:
        gen = solvers.get_general_stiffness_matrix(c, e, H.a11, H.a12, H.a21, H.a22, rule)
        wm = solvers.get_weighted_mass_matrix(c, e, u, H.phi, rule)
        rhs = solvers.get_right_hand_side(c, e, H.f_load, cubature_rule=rule)
        rhs0 = solvers.get_right_hand_side(c, e, H.f_load)
        minus_one = lambda r: -np.ones(r.shape[0])   # noqa: E731  (host result, uploaded)
        zero = lambda r: np.zeros(r.shape[0])        # noqa: E731
        ident = solvers.get_general_stiffness_matrix(c, e, minus_one, zero, zero,
                                                     minus_one, rule)
        dirichlet, neumann = np.vstack(b)[:40], np.vstack(b)[40:]
        x = u + 0.1 * c[:, 0]
        eta = indicators.compute_eta_r(x, c, e, dirichlet, neumann, L.f, L.g)
    finally:
        options.callbacks = 'host'
    H.csr_values_close(gen, orc.general_stiffness_coo(c, e, H.a11, H.a12, H.a21, H.a22,
                                                      w, p).tocsr(), RTOL)
    H.csr_values_close(wm, orc.weighted_mass_coo(c, e, u, H.phi, w, p).tocsr(), RTOL)
    ref = orc.load_vector_cubature(c, e, H.f_load, w, p)
    assert np.abs(rhs - ref).max() <= 1e-12 * np.abs(ref).max()
    ref0 = orc.load_vector_midpoint(c, e, H.f_load)
    assert np.abs(rhs0 - ref0).max() <= 1e-12 * np.abs(ref0).max()
    assert np.allclose(ident.toarray(), solvers.get_stiffness_matrix(c, e).toarray())
    ref_eta = orc.eta_r(x, c, e, dirichlet, neumann, L.f, L.g)
    assert np.abs(eta - ref_eta).max() <= 1e-12 * np.abs(ref_eta).max()


@pytest.mark.parametrize('n', [9, 12, 41, 60, 300])
def test_fans_of_every_size_class(api, n):
    """Valence 9..40: overflow storage + one thread per row; > 40: one block per
    row.  All operators that walk the fans."""
    ang = 2 * np.pi * np.arange(n) / n
    c = np.vstack([[0., 0.], np.column_stack([np.cos(ang), np.sin(ang)])])
    c[1:] *= (1 + 0.1 * np.sin(5 * ang))[:, None]
    e = np.column_stack([np.zeros(n, dtype=int), 1 + np.arange(n), 1 + (np.arange(n) + 1) % n])
    rim = np.column_stack([1 + np.arange(n), 1 + (np.arange(n) + 1) % n])
    H.csr_values_close(api[0].get_stiffness_matrix(c, e), orc.stiffness_coo(c, e).tocsr(), RTOL)
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.SMPLX1)
    H.csr_values_close(
        api[0].get_general_stiffness_matrix(c, e, H.a11, H.a12, H.a21, H.a22, (w, p)),
        orc.general_stiffness_coo(c, e, H.a11, H.a12, H.a21, H.a22, w, p).tocsr(), RTOL)
    assert np.array_equal(api[0].get_right_hand_side(c, e, H.f_load, (w, p)),
                          orc.load_vector_cubature(c, e, H.f_load, w, p))
    e2e, e2n, b2e = api[1].provide_geometric_data(e, [rim])
    o2e, o2n, ob2e = orc.geometric_data(e, [rim])
    assert np.array_equal(e2e, o2e) and np.array_equal(e2n, o2n) and np.array_equal(b2e[0], ob2e[0])
    x = H.u_nodal(c) + c[:, 1]
    assert np.array_equal(
        api[2].compute_eta_r(x, c, e, rim, np.zeros((0, 2), dtype=int), H.f_load, H.g_neu),
        orc.eta_r(x, c, e, rim, np.zeros((0, 2), dtype=int), H.f_load, H.g_neu))


def test_unstructured_delaunay_mesh(api):
    """A mesh that is NOT in NVB order: random points, Delaunay triangulation,
    valences 3..12 (rows with more than 8 incidences take the overflow pass),
    node numbering without any locality."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(7)
    pts = np.vstack([rng.random((4000, 2)),
                     [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]])
    tri = Delaunay(pts)
    e = tri.simplices.astype(np.int64)
    # counter-clockwise orientation, like every mesh of the reference
    d1, d2 = pts[e[:, 1]] - pts[e[:, 0]], pts[e[:, 2]] - pts[e[:, 0]]
    flip = (d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]) < 0
    e[flip] = e[flip][:, [0, 2, 1]]
    valence = np.bincount(e.reshape(-1))
    assert valence.max() > 8
    H.csr_values_close(api[0].get_stiffness_matrix(pts, e), orc.stiffness_coo(pts, e).tocsr(), RTOL)
    H.csr_values_close(api[0].get_mass_matrix(pts, e), orc.mass_coo(pts, e).tocsr(), RTOL)
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.DAYTAYLOR)
    u = H.u_nodal(pts)
    H.csr_values_close(api[0].get_weighted_mass_matrix(pts, e, u, H.phi, (w, p)),
                       orc.weighted_mass_coo(pts, e, u, H.phi, w, p).tocsr(), RTOL)
    assert np.array_equal(api[0].get_right_hand_side(pts, e, H.f_load, (w, p)),
                          orc.load_vector_cubature(pts, e, H.f_load, w, p))
    assert np.array_equal(api[0].get_right_hand_side(pts, e, H.f_load),
                          orc.load_vector_midpoint(pts, e, H.f_load))
    # boundary = edges that occur once, oriented like their element
    he = np.vstack([e[:, [0, 1]], e[:, [1, 2]], e[:, [2, 0]]])
    key = np.sort(he, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    bnd = he[cnt[np.asarray(inv).reshape(-1)] == 1]
    e2e, e2n, b2e = api[1].provide_geometric_data(e, [bnd])
    o2e, o2n, ob2e = orc.geometric_data(e, [bnd])
    assert np.array_equal(e2e, o2e) and np.array_equal(e2n, o2n) and np.array_equal(b2e[0], ob2e[0])
    x = u + 0.3 * pts[:, 0]
    none = np.zeros((0, 2), dtype=int)
    assert np.array_equal(api[2].compute_eta_r(x, pts, e, bnd, none, H.f_load, H.g_neu),
                          orc.eta_r(x, pts, e, bnd, none, H.f_load, H.g_neu))


def test_device_resident_api_and_plan_reuse():
    """DeviceMesh: torch CUDA tensors in, CSR triple of CUDA tensors out; with
    reuse=True a second operator on the same mesh only refreshes the record
    values (overflow rows included: Delaunay mesh with valence > 8)."""
    import torch
    from scipy.spatial import Delaunay
    from p1afempy_b200.device import DeviceMesh
    rng = np.random.default_rng(11)
    pts = np.vstack([rng.random((3000, 2)), [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]])
    e = Delaunay(pts).simplices.astype(np.int64)
    d1, d2 = pts[e[:, 1]] - pts[e[:, 0]], pts[e[:, 2]] - pts[e[:, 0]]
    flip = (d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]) < 0
    e[flip] = e[flip][:, [0, 2, 1]]
    assert np.bincount(e.reshape(-1)).max() > 8
    ct = torch.from_numpy(pts).cuda()
    et = torch.from_numpy(e.astype(np.int32)).cuda()
    ref_a = orc.stiffness_coo(pts, e).tocsr()
    ref_m = orc.mass_coo(pts, e).tocsr()
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.DAYTAYLOR)
    ref_g = orc.general_stiffness_coo(pts, e, H.a11, H.a12, H.a21, H.a22, w, p).tocsr()
    with DeviceMesh(ct, et, reuse=True) as mesh:
        ip, ix, a = mesh.stiffness_csr()
        assert ip.is_cuda and ix.dtype == torch.int32 and a.dtype == torch.float64
        assert np.array_equal(ip.cpu().numpy(), ref_a.indptr)
        assert np.array_equal(ix.cpu().numpy(), ref_a.indices)
        H.values_close_by_row(a.cpu().numpy(), ref_a, RTOL)
        plan_before = mesh._plan.value
        none_ip, none_ix, m = mesh.mass_csr(pattern=False)       # values only
        assert none_ip is None and none_ix is None and mesh._plan.value == plan_before
        H.values_close_by_row(m.cpu().numpy(), ref_m, RTOL)
        xq = mesh.quadrature_points(p)
        qs = [torch.from_numpy(np.stack([f(xq[q].cpu().numpy()) for q in range(len(w))])).cuda()
              for f in (H.a11, H.a12, H.a21, H.a22)]
        _, _, g = mesh.general_stiffness_csr(*qs, w, pattern=False)
        H.values_close_by_row(g.cpu().numpy(), ref_g, RTOL)
    with DeviceMesh(ct, et) as cold:     # quadrature sums inside the scatter
        gip, gix, g2 = cold.general_stiffness_csr(*qs, w, fused=True)
        assert np.array_equal(gix.cpu().numpy(), ref_g.indices)
        assert np.array_equal(g2.cpu().numpy(), g.cpu().numpy())      # same arithmetic
    with DeviceMesh(ct, et, reuse=True) as mesh:
        mesh.stiffness_csr()
        # and the vector / edge consumers work on the same (element-keeping) plan
        fc = torch.from_numpy(H.f_load(mesh.centroids().cpu().numpy())).cuda()
        b = mesh.load_vector_midpoint(fc)
        assert np.array_equal(b.cpu().numpy(), orc.load_vector_midpoint(pts, e, H.f_load))


def test_new_values_on_a_kept_plan_take_the_cold_path(mesh_case):
    """DeviceMesh(reuse=True): a second operator on the unchanged mesh is assembled by
    the cold path (faster than refreshing the plan's records in place) and gives the
    bits of the plan route; the plan stays for the vector / edge consumers."""
    import torch
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = mesh_case
    ct, et = torch.from_numpy(c).cuda(), torch.from_numpy(e.astype(np.int32)).cuda()
    with DeviceMesh(ct, et) as plain:
        ref_ip, ref_ix, ref_m = [t.cpu().numpy() for t in plain.mass_csr()]
    with DeviceMesh(ct, et, reuse=True) as mesh:
        mesh.stiffness_csr()
        plan_before = mesh._plan.value
        regular = int(mesh.assemble_cold(_lib.OP_MASS).status.cpu()[0]) == 0
        ip, ix, m = mesh.mass_csr()
        assert mesh._plan.value == plan_before and mesh._cold_refill == regular
        assert np.array_equal(ip.cpu().numpy(), ref_ip) and np.array_equal(ix.cpu().numpy(), ref_ix)
        assert np.array_equal(m.cpu().numpy(), ref_m)
        none_ip, none_ix, m2 = mesh.mass_csr(pattern=False)
        assert none_ip is None and none_ix is None and np.array_equal(m2.cpu().numpy(), ref_m)
        mesh._cold_refill = False                # the in-place refresh of the kept plan
        _, _, m3 = mesh.mass_csr(pattern=False)
        assert np.array_equal(m3.cpu().numpy(), ref_m)


def test_topology_plan_matches_element_keeping_plan():
    """P1_PLAN_TOPOLOGY (16-byte slots, no values, no row pointer) serves load
    vectors, edge numbering and eta_R with the same bits as the element-keeping
    plan; a mesh with a node of more than 8 elements falls back to that plan;
    asking such a plan for a matrix is an error at the C ABI."""
    import ctypes as C
    import torch
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import DeviceMesh
    pts, e, bnd = H.perturbed_square(4)
    bnds = [np.asarray(b, dtype=np.int64) for b in bnd]
    ct, et = torch.from_numpy(pts).cuda(), torch.from_numpy(e.astype(np.int32)).cuda()
    bt = [torch.from_numpy(b.astype(np.int32)).cuda() for b in bnds]
    out = {}
    for reuse in (False, True):
        with DeviceMesh(ct, et, reuse=reuse) as mesh:
            fc = torch.from_numpy(H.f_load(mesh.centroids().cpu().numpy())).cuda()
            gd = mesh.geometric_data(bt)
            assert mesh._plan_topology == (not reuse)
            b = mesh.load_vector_midpoint(fc)       # gathers through the existing plan
            out[reuse] = [b.cpu().numpy()] + [np.asarray(x.cpu().numpy()) for x in gd[:2]] + \
                         [x.cpu().numpy() for x in gd[2]]
            if not reuse:
                topo = mesh._plan
                rc = mesh._lib.p1_assemble_csr(topo, _lib.OP_STIFFNESS, None, None, None, None,
                                               None, None)
                assert rc == _lib.P1_EINVAL
                mesh.stiffness_csr()            # replaces the plan, still works
                assert not mesh._plan_topology
    for a, b in zip(out[False], out[True]):
        assert np.array_equal(a, b)
    # edge numbering on a one-shot plan builds only the twins it reads (half-edges
    # I -> J with J < I); an estimator call on the same plan completes them
    x = torch.from_numpy(H.u_nodal(pts)).cuda()
    dirichlet = torch.cat(bt)
    none = torch.zeros((0, 2), dtype=torch.int32, device='cuda')
    with DeviceMesh(ct, et) as mesh:
        fc = torch.from_numpy(H.f_load(mesh.centroids().cpu().numpy())).cuda()
        eta_first = mesh.eta_r(x, dirichlet, none, None, fc).cpu().numpy()
        gd = mesh.geometric_data(bt)            # twins by gather (P1_PLAN_SLOT_INDEX), all valid
        assert np.array_equal(gd[0].cpu().numpy(), out[True][1])
    with DeviceMesh(ct, et) as mesh:
        gd = mesh.geometric_data(bt)
        assert np.array_equal(gd[0].cpu().numpy(), out[True][1])
        assert np.array_equal(mesh.eta_r(x, dirichlet, none, None, fc).cpu().numpy(), eta_first)
        gd = mesh.geometric_data(bt)            # and the full twins serve the numbering
        assert np.array_equal(gd[0].cpu().numpy(), out[True][1])
    assert np.array_equal(eta_first, orc.eta_r(H.u_nodal(pts), pts, e, np.vstack(bnds),
                                               np.zeros((0, 2), dtype=int), H.f_load, H.g_neu))
    assert np.array_equal(out[False][0], orc.load_vector_midpoint(pts, e, H.f_load))
    o2e, o2n, ob2e = orc.geometric_data(e, bnds)
    assert np.array_equal(out[False][1], o2e) and np.array_equal(out[False][2], o2n)
    for got, want in zip(out[False][3:], ob2e):
        assert np.array_equal(got, want)
    # no plan at all: the plan-free load vectors (element code + value per slot)
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.DAYTAYLOR)
    with DeviceMesh(ct, et) as mesh:
        fc = torch.from_numpy(H.f_load(mesh.centroids().cpu().numpy())).cuda()
        b = mesh.load_vector_midpoint(fc)
        assert mesh._plan is None
        assert np.array_equal(b.cpu().numpy(), out[False][0])
        xq = mesh.quadrature_points(p)
        fq = torch.from_numpy(np.stack([H.f_load(xq[q].cpu().numpy())
                                        for q in range(len(w))])).cuda()
        bq = mesh.load_vector(fq, w, p)
        assert mesh._plan is None
        assert np.array_equal(bq.cpu().numpy(), orc.load_vector_cubature(pts, e, H.f_load, w, p))
    with DeviceMesh(ct, et, reuse=True) as mesh:
        assert np.array_equal(mesh.load_vector(fq, w, p).cpu().numpy(), bq.cpu().numpy())
        assert mesh._plan is not None
    # fans around one node: 12 elements -- 4 records beyond the 8 slots travel through
    # the spill lists of the plan-free load vectors and of the topology-only plan;
    # 1100 elements -- more than the lists hold: both are refused (P1_ERETRY) and the
    # element-keeping plan takes over
    for n, fits in ((12, True), (1100, False)):
        ang = 2 * np.pi * np.arange(n) / n
        fp = np.vstack([[0., 0.], np.c_[np.cos(ang), np.sin(ang)]])
        fe = np.array([[0, 1 + i, 1 + (i + 1) % n] for i in range(n)], dtype=np.int64)
        fb = np.array([[1 + i, 1 + (i + 1) % n] for i in range(n)], dtype=np.int64)
        xt = torch.from_numpy(H.u_nodal(fp)).cuda()
        bt12 = torch.from_numpy(fb.astype(np.int32)).cuda()
        o2e, o2n, ob2e = orc.geometric_data(fe, [fb])
        w1, p1 = cubature.resolve_rule(cubature.CubatureRuleEnum.SMPLX1)
        with DeviceMesh(torch.from_numpy(fp).cuda(),
                        torch.from_numpy(fe.astype(np.int32)).cuda()) as mesh:
            fc = torch.from_numpy(H.f_load(mesh.centroids().cpu().numpy())).cuda()
            b = mesh.load_vector_midpoint(fc)
            assert mesh._direct_fits == fits and (mesh._plan is None) == fits
            assert np.array_equal(b.cpu().numpy(), orc.load_vector_midpoint(fp, fe, H.f_load))
            xq = mesh.quadrature_points(p1)
            fq1 = torch.from_numpy(np.stack([H.f_load(xq[q].cpu().numpy())
                                             for q in range(len(w1))])).cuda()
            assert np.array_equal(mesh.load_vector(fq1, w1, p1).cpu().numpy(),
                                  orc.load_vector_cubature(fp, fe, H.f_load, w1, p1))
            mesh.close()
            eta = mesh.eta_r(xt, bt12, none, None, fc).cpu().numpy()
            assert mesh._plan_topology == fits and mesh._topology_fits == fits
            assert np.array_equal(eta, orc.eta_r(H.u_nodal(fp), fp, fe, fb,
                                                 np.zeros((0, 2), dtype=int), H.f_load, H.g_neu))
            gd = mesh.geometric_data([bt12])
            assert np.array_equal(gd[0].cpu().numpy(), o2e)
            assert np.array_equal(gd[1].cpu().numpy(), o2n)
            assert np.array_equal(gd[2][0].cpu().numpy(), ob2e[0])
            # load vectors through the plan that is there now (gather route)
            assert np.array_equal(mesh.load_vector_midpoint(fc).cpu().numpy(), b.cpu().numpy())
        with DeviceMesh(torch.from_numpy(fp).cuda(),
                        torch.from_numpy(fe.astype(np.int32)).cuda()) as mesh:
            gd = mesh.geometric_data([bt12])     # one-shot plan, backward twins only
            assert mesh._plan_topology == fits
            assert np.array_equal(gd[0].cpu().numpy(), o2e)
            assert np.array_equal(gd[2][0].cpu().numpy(), ob2e[0])


def test_spill_list_and_second_pass_agree(api, monkeypatch):
    """Rows with more than 8 records: their extra records travel through a spill
    list (tiled path and record path); when the list is too small the record path
    scans the elements a second time.  All routes give the same matrix (Delaunay
    mesh, valence up to 10+)."""
    from p1afempy_b200 import _lib, options
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(5)
    pts = np.vstack([rng.random((4000, 2)), [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]])
    e = Delaunay(pts).simplices.astype(np.int64)
    d1, d2 = pts[e[:, 1]] - pts[e[:, 0]], pts[e[:, 2]] - pts[e[:, 0]]
    flip = (d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]) < 0
    e[flip] = e[flip][:, [0, 2, 1]]
    assert np.bincount(e.reshape(-1)).max() > 8
    ref = orc.stiffness_coo(pts, e).tocsr()
    got = {}
    for cap, flags in ((None, 0), ('rec', _lib.PLAN_TILED), ('3', _lib.PLAN_TINY_SPILL)):
        monkeypatch.setattr(options, 'plan_flags', flags)
        a = api[0].get_stiffness_matrix(pts, e)
        assert np.array_equal(a.indptr, ref.indptr) and np.array_equal(a.indices, ref.indices)
        H.csr_values_close(a, ref, RTOL)
        m = api[0].get_mass_matrix(pts, e)
        H.csr_values_close(m, orc.mass_coo(pts, e).tocsr(), RTOL)
        got[cap] = a.data.copy()
    # off-diagonals identical; diagonals of high-valence rows may differ in the last
    # bits (their summation order is the arrival order of the records)
    assert np.abs(got[None] - got['3']).max() <= 1e-13 * np.abs(ref.data).max()
    assert np.abs(got['rec'] - got['3']).max() <= 1e-13 * np.abs(ref.data).max()


def test_unaligned_element_array():
    """The element array as a view at a 4-byte offset (no 128-bit loads possible):
    every kernel takes its scalar path and the results do not change."""
    import torch
    from p1afempy_b200.device import DeviceMesh
    pts, e, bnd = H.perturbed_square(4)
    m = e.shape[0]
    assert m % 4 == 0
    e = e[:m - 3]                      # and a tail that is not a multiple of 4 triangles
    keep = np.unique(e)
    assert keep.size == pts.shape[0]   # (all nodes still used)
    ct = torch.from_numpy(pts).cuda()
    aligned = torch.from_numpy(e.astype(np.int32)).cuda()
    buf = torch.empty(3 * e.shape[0] + 1, dtype=torch.int32, device='cuda')
    buf[1:] = aligned.reshape(-1)
    shifted = buf[1:].view(-1, 3)
    assert shifted.data_ptr() % 16 != 0 and shifted.is_contiguous()
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.DAYTAYLOR)
    out = []
    for et in (aligned, shifted):
        with DeviceMesh(ct, et) as mesh:
            ip, ix, a = mesh.stiffness_csr()
            fc = torch.from_numpy(H.f_load(mesh.centroids().cpu().numpy())).cuda()
            xq = mesh.quadrature_points(p)
            fq = torch.from_numpy(np.stack([H.f_load(xq[q].cpu().numpy())
                                            for q in range(len(w))])).cuda()
            out.append([t.cpu().numpy() for t in
                        (ip, ix, a, mesh.load_vector_midpoint(fc), mesh.load_vector(fq, w, p),
                         mesh.mass_csr()[2], mesh.weighted_mass_csr(fq, w, p)[2],
                         mesh.general_stiffness_csr(fq, fq + 1., fq * 2., fq - 1., w)[2])])
    for x, y in zip(*out):
        assert np.array_equal(x, y)
    ref = orc.stiffness_coo(pts, e).tocsr()
    assert np.array_equal(out[1][1], ref.indices)
    assert np.array_equal(out[1][3], orc.load_vector_midpoint(pts, e, H.f_load))


def test_slim_plans_match_record_plans(api, mesh_case, monkeypatch):
    """Multi-GPU plans for stiffness / mass keep 8-byte slots and recompute the
    entries while filling (P1_PLAN_FORCE_SLIM forces that on one GPU): same pattern, same
    off-diagonals bit for bit, diagonals within the usual rounding."""
    from p1afempy_b200 import _lib, options
    c, e, _ = mesh_case
    monkeypatch.setattr(options, 'plan_flags', 0)
    a0 = api[0].get_stiffness_matrix(c, e)
    m0 = api[0].get_mass_matrix(c, e)
    monkeypatch.setattr(options, 'plan_flags', _lib.PLAN_FORCE_SLIM)
    a1 = api[0].get_stiffness_matrix(c, e)
    m1 = api[0].get_mass_matrix(c, e)
    for x, y in ((a0, a1), (m0, m1)):
        assert np.array_equal(x.indptr, y.indptr) and np.array_equal(x.indices, y.indices)
        off = x.indices != np.repeat(np.arange(x.shape[0]), np.diff(x.indptr))
        assert np.array_equal(x.data[off], y.data[off])
        H.csr_values_close(y, x, RTOL)
    H.csr_values_close(a1, orc.stiffness_coo(c, e).tocsr(), RTOL)


@pytest.mark.parametrize('rule', ['MIDPOINT', 'DAYTAYLOR'])
def test_tiled_plans_match_record_plans(api, mesh_case, rule, monkeypatch):
    """P1_PLAN_TILED (tile-interior rows finished in shared memory, the others
    through records): every matrix is bit for bit the record path's -- both build
    a row in the same arithmetic order."""
    from p1afempy_b200 import _lib, options
    c, e, _ = mesh_case
    r = cubature.CubatureRuleEnum[rule]
    u = H.u_nodal(c)
    out = []
    for flags in (0, _lib.PLAN_TILED):
        monkeypatch.setattr(options, 'plan_flags', flags)
        out.append([api[0].get_stiffness_matrix(c, e), api[0].get_mass_matrix(c, e),
                    api[0].get_general_stiffness_matrix(c, e, H.a11, H.a12, H.a21, H.a22, r),
                    api[0].get_weighted_mass_matrix(c, e, u, H.phi, r)])
    for x, y in zip(*out):
        assert x.shape == y.shape
        assert np.array_equal(x.indptr, y.indptr) and np.array_equal(x.indices, y.indices)
        assert np.array_equal(x.data, y.data)
    H.csr_values_close(out[1][0], orc.stiffness_coo(c, e).tocsr(), RTOL)


def test_tiled_plan_irregular_meshes(api, monkeypatch):
    """The tiled path on meshes that defeat its shortcuts: random element order (no
    row completes inside a tile), a non-manifold vertex (closed-fan shortcut
    contradicted -> exact plan), an unused node, 7 triangles (a partial quad)."""
    from p1afempy_b200 import _lib, options
    monkeypatch.setattr(options, 'plan_flags', _lib.PLAN_TILED)
    c, e, _ = H.perturbed_square(5)
    perm = np.random.default_rng(3).permutation(e.shape[0])
    for ee in (e[perm], e[:e.shape[0] - 1], e[perm][:4099]):
        cc = c
        a = api[0].get_stiffness_matrix(cc, ee)
        ref = orc.stiffness_coo(cc, ee).tocsr()
        assert a.shape == ref.shape
        assert np.array_equal(a.indptr, ref.indptr) and np.array_equal(a.indices, ref.indices)
        H.csr_values_close(a, ref, RTOL)
    # bow tie: two closed fans that share one vertex
    c2 = np.vstack([c, c + [2., 0.]])
    e2 = np.vstack([e, e + c.shape[0]])
    centre = int(np.argmin(np.abs(c - 0.5).sum(axis=1)))
    e2[e2 == centre + c.shape[0]] = centre
    a = api[0].get_stiffness_matrix(c2, e2)
    ref = orc.stiffness_coo(c2, e2).tocsr()
    assert np.array_equal(a.indptr, ref.indptr) and np.array_equal(a.indices, ref.indices)
    H.csr_values_close(a, ref, RTOL)


def test_cold_assembly_without_host_sync(mesh_case):
    """p1_assemble_cold (no host synchronisation, outputs with a capacity) gives the
    plan route's CSR bit for bit: all rows, and the shares of 3 emulated ranks."""
    import torch
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = mesh_case
    ct, et = torch.from_numpy(c).cuda(), torch.from_numpy(e.astype(np.int32)).cuda()
    for op in (_lib.OP_STIFFNESS, _lib.OP_MASS):
        with DeviceMesh(ct, et) as mesh:
            ref = [t.cpu().numpy() for t in mesh.assemble(op)]
            nnz = mesh.nnz
            cold = mesh.assemble_cold(op)
            assert cold.data.numel() >= nnz
            if int(cold.status[0].item()) == 0:     # (else: the plan route, via wait())
                assert int(cold.indptr[-1].item()) == nnz == int(cold.status[6].item())
            got = [t.cpu().numpy() for t in cold.wait()]
            assert mesh.nnz == nnz
        for x, y in zip(ref, got):
            assert x.dtype == y.dtype and np.array_equal(x, y)
        for rank in range(3):
            with DeviceMesh(ct, et, interleave=(3, rank, 4)) as mesh:
                ref = [t.cpu().numpy() for t in mesh.assemble(op)]
                got = [t.cpu().numpy() for t in mesh.assemble_cold(op).wait()]
            for x, y in zip(ref, got):
                assert np.array_equal(x, y)


def test_cold_assembly_replayed_as_a_graph(api):
    """A cold assembly of a small mesh into the same arrays is captured on its second
    call and replayed as one CUDA graph afterwards (plan.cu, COLD_GRAPH_MAX): every
    call must give the plan route's CSR, also after the mesh's CONTENT changed in place,
    after more argument sets than the cache holds, on a side stream, and after a trim."""
    import torch
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import DeviceMesh, context
    c, e, _ = H.perturbed_square(5)
    ct, et = torch.from_numpy(c).cuda(), torch.from_numpy(e.astype(np.int32)).cuda()
    lib = _lib.load()

    def expect(op):
        with DeviceMesh(ct.clone(), et.clone()) as m:
            return [t.cpu().numpy() for t in m.assemble(op)]

    def same(cold, ref):
        assert int(cold.status[0].item()) == 0
        for x, y in zip(ref, cold.wait()):
            assert np.array_equal(x, y.cpu().numpy())

    with DeviceMesh(ct, et) as mesh:
        for op in (_lib.OP_STIFFNESS, _lib.OP_MASS):
            ref = expect(op)
            cold = None
            for _ in range(4):                  # plain, capture, replay, replay
                cold = mesh.assemble_cold(op, out=cold)
                same(cold, ref)
                cold._done = None
                cold.data.fill_(-1.)            # a replay has to write everything again
                cold.indptr.zero_()
        # the coordinates move, same arrays: the graph reads the new content
        ct.mul_(1.5)
        ref = expect(_lib.OP_MASS)
        cold = mesh.assemble_cold(_lib.OP_MASS, out=cold)
        same(cold, ref)
        # more argument sets than graph slots; the first set comes back afterwards
        ref = expect(_lib.OP_STIFFNESS)
        sets = [None] * 11
        for _ in range(3):
            for i in range(len(sets)):
                sets[i] = mesh.assemble_cold(_lib.OP_STIFFNESS, out=sets[i])
                same(sets[i], ref)
                sets[i]._done = None
        # another stream
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            cold = None
            for _ in range(3):
                cold = mesh.assemble_cold(_lib.OP_STIFFNESS, out=cold)
                same(cold, ref)
                cold._done = None
        side.synchronize()
        # a trim drops the graphs and their scratch; the next call starts over
        assert lib.p1_ctx_trim(context()[0]) == 0
        for _ in range(3):
            sets[0] = mesh.assemble_cold(_lib.OP_STIFFNESS, out=sets[0])
            same(sets[0], ref)
            sets[0]._done = None
        with pytest.raises(ValueError):
            with DeviceMesh(*[torch.from_numpy(a).cuda() for a in
                              (H.perturbed_square(3)[0],
                               H.perturbed_square(3)[1].astype(np.int32))]) as other:
                other.assemble_cold(_lib.OP_STIFFNESS, out=sets[0])


def test_cold_assembly_from_two_threads(api):
    """Two host threads share the context (one per device), each on its own stream and
    its own output arrays: the entry points serialise on the context's lock, the cached
    graphs are keyed by stream and arrays, and every result is the plan route's."""
    import threading
    import torch
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = H.perturbed_square(5)
    ct, et = torch.from_numpy(c).cuda(), torch.from_numpy(e.astype(np.int32)).cuda()
    with DeviceMesh(ct, et) as m:
        refs = {op: [t.cpu().numpy() for t in m.assemble(op)]
                for op in (_lib.OP_STIFFNESS, _lib.OP_MASS)}
    torch.cuda.synchronize()
    errors = []

    def work(op):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream), DeviceMesh(ct, et) as mesh:
                cold = None
                for _ in range(40):
                    cold = mesh.assemble_cold(op, out=cold)
                    got = [t.cpu().numpy() for t in cold.wait()]
                    cold._done = None
                    for x, y in zip(refs[op], got):
                        if not np.array_equal(x, y):
                            raise AssertionError('mismatch in thread of op %d' % op)
        except Exception as ex:       # noqa: BLE001  (reported in the main thread)
            errors.append(ex)
    threads = [threading.Thread(target=work, args=(op,)) for op in refs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_cold_assembly_falls_back_to_the_plan_route(api):
    """Status words instead of return codes: a node with more than 8 elements
    (P1_STATUS_SLOWPATH), a non-manifold vertex (P1_STATUS_RETRY) and an index out
    of range are found when the result is waited for."""
    import torch
    from scipy.spatial import Delaunay
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import DeviceMesh
    rng = np.random.default_rng(5)
    pts = np.vstack([rng.random((3000, 2)), [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]])
    e = Delaunay(pts).simplices.astype(np.int64)
    d1, d2 = pts[e[:, 1]] - pts[e[:, 0]], pts[e[:, 2]] - pts[e[:, 0]]
    flip = (d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0]) < 0
    e[flip] = e[flip][:, [0, 2, 1]]
    assert np.bincount(e.reshape(-1)).max() > 8
    c, e5, b5 = H.perturbed_square(4)
    c2 = np.vstack([c, c + [2., 0.]])
    e2 = np.vstack([e5, e5 + c.shape[0]])
    inner = np.setdiff1d(np.flatnonzero(np.bincount(e5.reshape(-1)) == 4), np.vstack(b5))
    centre = int(inner[len(inner) // 2])
    e2[e2 == centre + c.shape[0]] = centre           # bow tie: two closed fans of 4 (a valid row)
    c0, e0, _ = meshgen.criss_square()
    e0 = np.vstack([e0, e0])     # every element twice: the centre's fan "closes" with repeats
    for cc, ee, bit in ((pts, e, _lib.STATUS_SLOWPATH), (c0, e0, _lib.STATUS_RETRY),
                        (c2, e2, 0)):
        with DeviceMesh(cc, ee) as mesh:
            cold = mesh.assemble_cold(_lib.OP_STIFFNESS)
            assert int(cold.status[0].item()) == bit
            ip, ix, data = [t.cpu().numpy() for t in cold.wait()]
        ref = orc.stiffness_coo(cc, ee).tocsr()
        assert np.array_equal(ip, ref.indptr) and np.array_equal(ix, ref.indices)
        assert np.abs(data - ref.data).max() <= RTOL * np.abs(ref.data).max()
    bad = e5.copy()
    bad[3, 1] = c.shape[0] + 7
    mesh = DeviceMesh(c, bad, check=False)
    with pytest.raises(IndexError):
        mesh.assemble_cold(_lib.OP_STIFFNESS).wait()
    mesh.close()
    # nnz above the capacity 2 N + 3 M: two triangles that share ONE vertex (5 nodes, 6
    # edges, all on the boundary: nnz = 17 > 16) -- reported, nothing written, plan route
    ct = np.array([[0., 0.], [1., 0.], [0., 1.], [-1., 0.], [0., -1.]])
    et = np.array([[0, 1, 2], [0, 3, 4]])
    with DeviceMesh(ct, et) as mesh:
        assert mesh.cold_capacity() == 16
        cold = mesh.assemble_cold(_lib.OP_STIFFNESS)
        assert int(cold.status[0].item()) == _lib.STATUS_SLOWPATH
        assert int(cold.status[6].item()) == 17
        ip, ix, data = [t.cpu().numpy() for t in cold.wait()]
    ref = orc.stiffness_coo(ct, et).tocsr()
    assert ref.nnz == 17 and np.array_equal(ip, ref.indptr) and np.array_equal(ix, ref.indices)
    assert np.abs(data - ref.data).max() <= RTOL * np.abs(ref.data).max()


def test_to_scipy_int64_branch(api, monkeypatch):
    """scipy switches the CSR index dtype to int64 when max(9 M, N) > 2^31 - 1
    (scipy/sparse/_coo.py:419; the 256M-triangle mesh of the sweep).  The same code
    path on a small mesh: the rule is forced, the device widens its int32 arrays
    (p1_widen_indices) and the matrix still equals the reference's."""
    import torch
    from p1afempy_b200 import device
    assert device.scipy_index_dtype(268435456, 134234113, 134234113) == np.int64
    assert device.scipy_index_dtype(67108864, 33562625, 33562625) == np.int32
    monkeypatch.setattr(device, 'scipy_index_dtype', lambda *a: np.int64)
    widened = []
    real = device.DeviceMesh.widen_indices
    monkeypatch.setattr(device.DeviceMesh, 'widen_indices',
                        lambda self, t: widened.append(real(self, t)) or widened[-1])
    c, e, _ = H.perturbed_square(4)
    for fn, ref in ((api[0].get_stiffness_matrix, orc.stiffness_coo(c, e).tocsr()),
                    (api[0].get_mass_matrix, orc.mass_coo(c, e).tocsr())):
        del widened[:]
        a = fn(c, e)
        # (scipy's constructor narrows int64 arrays whose values fit int32 again; at
        # 256M triangles they do not: tests/test_gpu_fullsize.py)
        assert len(widened) == 2 and all(t.dtype == torch.int64 for t in widened)
        assert np.array_equal(widened[0].cpu().numpy(), ref.indptr)
        assert np.array_equal(widened[1].cpu().numpy(), ref.indices)
        assert a.has_sorted_indices and a.shape == ref.shape
        assert np.array_equal(a.indptr, ref.indptr) and np.array_equal(a.indices, ref.indices)
        H.values_close_by_row(a.data, ref, RTOL)


def test_integration_md_stub_is_real(api):
    """The ctypes stub printed in INTEGRATION.md (what a maintainer would paste into
    the reference's solvers.py) runs as written and returns reference(...).tocsr()."""
    import re
    from p1afempy_b200 import _lib
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(repo, 'INTEGRATION.md')).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if 'def get_stiffness_matrix' in b]
    assert len(stub) == 1
    code = stub[0].replace('C.CDLL("libp1b200.so")', 'C.CDLL(%r)' % _lib.LIB_PATH)
    ns = {}
    exec(compile(code, 'INTEGRATION.md', 'exec'), ns)
    for c, e, _ in (H.perturbed_square(3), meshgen.l_shape()):
        a = ns['get_stiffness_matrix'](c, e)
        ref = orc.stiffness_coo(c, e).tocsr()
        assert a.shape == ref.shape
        assert np.array_equal(a.indptr, ref.indptr) and np.array_equal(a.indices, ref.indices)
        H.csr_values_close(a, ref, RTOL)


def test_empty_and_tiny_inputs(api):
    """Empty element arrays (the reference: scipy cannot infer a shape ->
    ValueError; explicit-shape builders return empty matrices) and a single
    triangle."""
    solvers, mesh, indicators = api
    c = np.array([[0., 0.], [1., 0.], [0., 1.]])
    none = np.zeros((0, 3), dtype=np.int64)
    with pytest.raises(ValueError):
        solvers.get_stiffness_matrix(c, none)
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.MIDPOINT)
    g = solvers.get_general_stiffness_matrix(c, none, H.a11, H.a12, H.a21, H.a22, (w, p))
    assert g.shape == (3, 3) and g.nnz == 0
    assert np.array_equal(solvers.get_right_hand_side(c, none, H.f_load), np.zeros(3))
    assert np.array_equal(solvers.get_right_hand_side(c, none, H.f_load, (w, p)), np.zeros(3))
    one = np.array([[0, 1, 2]])
    H.csr_values_close(solvers.get_stiffness_matrix(c, one), orc.stiffness_coo(c, one).tocsr(), RTOL)
    H.csr_values_close(solvers.get_mass_matrix(c, one), orc.mass_coo(c, one).tocsr(), RTOL)
    assert np.array_equal(solvers.get_right_hand_side(c, one, H.f_load),
                          orc.load_vector_midpoint(c, one, H.f_load))
    edges = np.array([[0, 1], [1, 2], [2, 0]])
    e2e, e2n, b2e = mesh.provide_geometric_data(one, [edges])
    o2e, o2n, ob2e = orc.geometric_data(one, [edges])
    assert np.array_equal(e2e, o2e) and np.array_equal(e2n, o2n) and np.array_equal(b2e[0], ob2e[0])
    x = np.array([0., 1., 2.])
    assert np.array_equal(
        indicators.compute_eta_r(x, c, one, edges[:2], edges[2:], H.f_load, H.g_neu),
        orc.eta_r(x, c, one, edges[:2], edges[2:], H.f_load, H.g_neu))
    # a clockwise triangle: negative area propagates exactly like in numpy
    cw = np.array([[0, 2, 1]])
    H.csr_values_close(solvers.get_mass_matrix(c, cw), orc.mass_coo(c, cw).tocsr(), RTOL)
    assert solvers.get_mass_matrix(c, cw).data.max() < 0


def test_nonlinear_cousins(api, mesh_case):
    """load vector of f(u_h) (bit-exact) and the
    integral of f(u_h) (fixed-order tree vs numpy's pairwise sum: 1e-12)."""
    c, e, _ = mesh_case
    u = H.u_nodal(c)
    for rule in ('MIDPOINT', 'DAYTAYLOR'):
        r = cubature.CubatureRuleEnum[rule]
        w, p = cubature.resolve_rule(r)
        got = api[0].get_load_vector_of_composition_nonlinear_with_fem(H.f_nl, u, c, e, r)
        assert np.array_equal(got, orc.load_vector_composition(H.f_nl, u, c, e, w, p))
        val = api[0].integrate_composition_nonlinear_with_fem(H.f_nl, u, c, e, r)
        ref = orc.integrate_composition(H.f_nl, u, c, e, w, p)
        assert abs(val - ref) <= 1e-12 * abs(ref)


def test_nonlinear_cousins_against_reference_fixture(api):
    g = load('graded')
    c, e, u = g['coordinates'], g['elements'], g['u']
    rule = (g['rule_DAYTAYLOR_w'], g['rule_DAYTAYLOR_p'])
    assert np.array_equal(
        api[0].get_load_vector_of_composition_nonlinear_with_fem(H.f_nl, u, c, e, rule),
        g['nl_load_DAYTAYLOR'])
    val = api[0].integrate_composition_nonlinear_with_fem(H.f_nl, u, c, e, rule)
    assert abs(val - float(g['nl_integral_DAYTAYLOR'])) <= 1e-12 * abs(val)


def test_adaptive_lshape_loop(api):
    """Config C4 at test size: L-shaped domain, nodal
    interpolant of r^(2/3) sin(2 theta/3) about the re-entrant corner, eta_R on
    the GPU drives Doerfler marking (theta = 0.5, stable descending sort) and
    NVB refinement; at every step eta_R equals the oracle bit for bit and the
    mesh grades towards the corner."""
    solvers, mesh, indicators = api
    c, e, b = meshgen.refine_uniform(*meshgen.l_shape(), times=3)
    dirichlet_of = lambda bs: np.vstack(bs)               # noqa: E731
    none = np.zeros((0, 2), dtype=int)

    def singular(r):
        x, y = r[:, 0] - 1., r[:, 1] - 1.
        theta = np.mod(np.arctan2(y, x) - np.pi / 2., 2. * np.pi)   # 0 along the edge x=1, y>1
        return np.hypot(x, y) ** (2. / 3.) * np.sin(2. * theta / 3.)
    zero = lambda r: np.zeros(r.shape[0])                 # noqa: E731
    sizes = []
    for _ in range(7):
        x = singular(c)
        eta = indicators.compute_eta_r(x, c, e, dirichlet_of(b), none, zero, zero)
        assert np.array_equal(eta, orc.eta_r(x, c, e, dirichlet_of(b), none, zero, zero))
        order = np.argsort(-eta, kind='stable')
        csum = np.cumsum(eta[order])
        marked = order[:int(np.searchsorted(csum, 0.5 * csum[-1])) + 1]
        sizes.append(e.shape[0])
        c, e, b = meshgen.refine_nvb(c, e, marked, b)
    assert sizes[-1] > sizes[0]
    # the estimator concentrates the refinement at the re-entrant corner
    cent = (c[e[:, 0]] + c[e[:, 1]] + c[e[:, 2]]) / 3.
    area = np.abs(orc.signed_area(c, e))
    near = np.hypot(cent[:, 0] - 1., cent[:, 1] - 1.) < 0.1
    assert area[near].min() < 0.01 * area[~near].max()
    a = solvers.get_stiffness_matrix(c, e)
    H.csr_values_close(a, orc.stiffness_coo(c, e).tocsr(), RTOL)


def test_random_element_soups(api):
    """Differential fuzzing of the row machinery: random triples over a few
    nodes (non-manifold, repeated vertices, duplicated triangles, both
    orientations) -- the reference just sums triplets, so must we.  The mass
    matrix has no division by the area, so every value stays finite."""
    rng = np.random.default_rng(2024)
    for trial in range(120):
        n = int(rng.integers(3, 14))
        m = int(rng.integers(1, 40))
        c = rng.random((n, 2))
        e = rng.integers(0, n, size=(m, 3))
        if trial % 3 == 0:                       # proper triangles only
            e = np.array([rng.choice(n, size=3, replace=False) for _ in range(m)])
        ref = orc.mass_coo(c, e).tocsr()
        got = api[0].get_mass_matrix(c, e)
        assert got.shape == ref.shape, trial
        assert np.array_equal(got.indptr, ref.indptr), trial
        assert np.array_equal(got.indices, ref.indices), trial
        # both orientations cancel: measure against the size of the contributions
        scale = np.abs(orc.mass_triplets(c, e)[2]).max()
        assert np.abs(got.data - ref.data).max() <= 1e-12 * scale, trial
        b = api[0].get_right_hand_side(c, e, H.f_load)
        assert np.array_equal(b, orc.load_vector_midpoint(c, e, H.f_load)), trial


def test_solve_laplace_matlab_golden(api):
    """The reference's own end-to-end tests on the drop-in functions:
    test_solver.py:20-57 (x and energy of MATLAB's p1afem after 3 NVB
    refinements) and test_indicators.py:12-55 (eta_R after 2)."""
    solvers, _, indicators = api
    g = load('laplace_example')
    x3, energy3 = solvers.solve_laplace(
        coordinates=g['coordinates3'], elements=g['elements3'], dirichlet=g['dirichlet3'],
        neumann=g['neumann3'], f=L.f, g=L.g, uD=L.uD)
    assert np.allclose(x3, g['x_matlab'])
    assert np.isclose(energy3, g['energy_matlab'])
    assert np.allclose(x3, g['x3_reference'], rtol=1e-12, atol=1e-14)
    x2, _ = solvers.solve_laplace(
        coordinates=g['coordinates'], elements=g['elements'], dirichlet=g['dirichlet'],
        neumann=g['neumann'], f=L.f, g=L.g, uD=L.uD)
    eta = indicators.compute_eta_r(x2, g['coordinates'], g['elements'], g['dirichlet'],
                                   g['neumann'], L.f, L.g)
    assert np.allclose(eta, g['eta_matlab'])


def test_weighted_mass_closed_form(api):
    """test_get_weighted_mass_matrix.py:20-96 of the reference on the drop-in:
    u^T M(phi(u)) u for a pyramid function and a degree-4 polynomial phi equals
    a closed form on every refined mesh (needs a rule exact to degree 6)."""
    coefficients = [34.3, 12.4, 1.4, 0.45, 0.332]
    u_star = 12.
    c = np.array([[-1., -1.], [1., -1.], [1., 1.], [-1., 1.], [0., 0.]])
    e = np.array([[0, 4, 3], [0, 1, 4], [1, 2, 4], [4, 2, 3]])
    b = [np.array([[0, 1], [1, 2], [2, 3], [3, 0]])]
    u = np.array([0., 0., 0., 0., u_star])

    def phi(x):
        out = np.zeros(x.shape[0])
        for k, a in enumerate(coefficients):
            out += a * x ** k
        return out
    exact = 8. * u_star ** 2 * sum(a * u_star ** k / ((k + 3.) * (k + 4.))
                                   for k, a in enumerate(coefficients))
    for _ in range(5):
        # refineNVB with to_embed: new nodes take the mean of their edge's ends
        e2e, e2n, _ = meshgen.number_edges(e, b)
        u = np.concatenate([u, (u[e2n[:, 0]] + u[e2n[:, 1]]) / 2.])
        c, e, b = meshgen.refine_nvb(c, e, np.arange(e.shape[0]), b)
        m = api[0].get_weighted_mass_matrix(coordinates=c, elements=e, current_iterate=u,
                                            phi=phi,
                                            cubature_rule=cubature.CubatureRuleEnum.DAYTAYLOR)
        assert abs(u.dot(m.dot(u)) - exact) < 5e-8 * abs(exact)
