"""Pins the oracle: every restated function must equal the REAL reference
(imported from /root/reference) bit for bit on generic inputs.  Skipped on the
GPU box, where the reference is not mounted; there the committed golden
fixtures (test_oracle_golden.py) pin it instead."""
import numpy as np
import pytest

from oracle import p1_oracle as orc
from oracle import ref_import
from p1afempy_b200 import cubature
import helpers as H

pytestmark = pytest.mark.skipif(not ref_import.available(),
                                reason='reference not mounted')

RULES = [cubature.CubatureRuleEnum.MIDPOINT, cubature.CubatureRuleEnum.SMPLX1,
         cubature.CubatureRuleEnum.DAYTAYLOR]


@pytest.fixture(scope='module')
def ref():
    return ref_import.load()


@pytest.fixture(scope='module', params=['perturbed', 'graded'])
def mesh(request):
    if request.param == 'perturbed':
        return H.perturbed_square(4)
    return H.graded_square()


def _same_coo(a, b):
    assert a.shape == b.shape
    assert np.array_equal(a.row, b.row) and np.array_equal(a.col, b.col)
    assert np.array_equal(a.data, b.data)          # bit-exact
    x, y = a.tocsr(), b.tocsr()
    assert x.indices.dtype == y.indices.dtype
    assert np.array_equal(x.indptr, y.indptr)
    assert np.array_equal(x.indices, y.indices)
    assert np.array_equal(x.data, y.data)


def test_geometry(ref, mesh):
    c, e, _ = mesh
    p, q = orc.edge_vectors(c, e)
    rp, rq = ref.mesh.get_directional_vectors(c, e)
    assert np.array_equal(p, rp) and np.array_equal(q, rq)
    assert np.array_equal(orc.signed_area(c, e), ref.mesh.get_area(c, e))


def test_stiffness(ref, mesh):
    c, e, _ = mesh
    _same_coo(orc.stiffness_coo(c, e), ref.solvers.get_stiffness_matrix(c, e))


def test_mass(ref, mesh):
    c, e, _ = mesh
    i, j, d = orc.mass_triplets(c, e)
    ri, rj, rd = ref.solvers.get_mass_matrix_elements(c, e)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    assert np.array_equal(d, rd)
    _same_coo(orc.mass_coo(c, e), ref.solvers.get_mass_matrix(c, e))


@pytest.mark.parametrize('rule', RULES)
def test_general_stiffness(ref, mesh, rule):
    c, e, _ = mesh
    w, p = cubature.resolve_rule(rule)
    mine = orc.general_stiffness_coo(c, e, H.a11, H.a12, H.a21, H.a22, w, p)
    theirs = ref.solvers.get_general_stiffness_matrix(
        coordinates=c, elements=e, a_11=H.a11, a_12=H.a12, a_21=H.a21,
        a_22=H.a22, cubature_rule=rule)
    _same_coo(mine, theirs)


@pytest.mark.parametrize('rule', RULES)
def test_weighted_mass(ref, mesh, rule):
    c, e, _ = mesh
    u = H.u_nodal(c)
    w, p = cubature.resolve_rule(rule)
    mine = orc.weighted_mass_coo(c, e, u, H.phi, w, p)
    theirs = ref.solvers.get_weighted_mass_matrix(
        coordinates=c, elements=e, current_iterate=u, phi=H.phi,
        cubature_rule=rule)
    _same_coo(mine, theirs)


@pytest.mark.parametrize('rule', [None] + RULES)
def test_load_vector(ref, mesh, rule):
    c, e, _ = mesh
    theirs = ref.solvers.get_right_hand_side(c, e, H.f_load,
                                             cubature_rule=rule)
    if rule is None:
        mine = orc.load_vector_midpoint(c, e, H.f_load)
    else:
        w, p = cubature.resolve_rule(rule)
        mine = orc.load_vector_cubature(c, e, H.f_load, w, p)
    assert np.array_equal(mine, theirs)


@pytest.mark.parametrize('rule', RULES)
def test_nonlinear_cousins(ref, mesh, rule):
    c, e, _ = mesh
    u = H.u_nodal(c)
    w, p = cubature.resolve_rule(rule)
    assert np.array_equal(
        orc.load_vector_composition(H.f_nl, u, c, e, w, p),
        ref.solvers.get_load_vector_of_composition_nonlinear_with_fem(
            f=H.f_nl, u=u, coordinates=c, elements=e, cubature_rule=rule))
    assert orc.integrate_composition(H.f_nl, u, c, e, w, p) == \
        ref.solvers.integrate_composition_nonlinear_with_fem(
            f=H.f_nl, u=u, coordinates=c, elements=e, cubature_rule=rule)


def _split_boundary(b):
    """dirichlet = first ~60 % of the edges, neumann = the rest."""
    allb = np.vstack(b)
    cut = (3 * allb.shape[0]) // 5
    return allb[:cut], allb[cut:]


def test_geometric_data(ref, mesh):
    _, e, b = mesh
    parts = list(_split_boundary(b))
    mine = orc.geometric_data(e, parts)
    theirs = ref.mesh.provide_geometric_data(e, parts)
    assert np.array_equal(mine[0], theirs[0]) and mine[0].dtype == theirs[0].dtype
    assert np.array_equal(mine[1], theirs[1])
    assert len(mine[2]) == len(theirs[2])
    for x, y in zip(mine[2], theirs[2]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize('with_neumann', [True, False])
def test_eta_r(ref, mesh, with_neumann):
    c, e, b = mesh
    if with_neumann:
        dirichlet, neumann = _split_boundary(b)
    else:
        dirichlet, neumann = np.vstack(b), np.zeros((0, 2), dtype=int)
    x = H.u_nodal(c) + 0.1 * c[:, 0]
    mine = orc.eta_r(x, c, e, dirichlet, neumann, H.f_load, H.g_neu)
    theirs = ref.indicators.compute_eta_r(x, c, e, dirichlet, neumann,
                                          H.f_load, H.g_neu)
    assert np.array_equal(mine, theirs)


def test_incomplete_boundary_raises(ref, mesh):
    _, e, b = mesh
    partial = [np.vstack(b)[:-3]]
    with pytest.raises(ValueError):
        ref.mesh.provide_geometric_data(e, partial)
    with pytest.raises(ValueError):
        orc.geometric_data(e, partial)
