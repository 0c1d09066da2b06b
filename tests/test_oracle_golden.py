"""The oracle against the committed golden fixtures (produced by the REAL
reference, tests/golden/make_golden.py) and against the MATLAB vectors /
known answers the reference's own tests hold.  Runs without the reference."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from oracle import p1_oracle as orc
import helpers as H
import helpers_laplace as L


def load(name):
    return np.load(os.path.join(H.GOLDEN, name + '.npz'))


def csr_of(g, prefix):
    return csr_matrix((g[prefix + '_data'], g[prefix + '_indices'],
                       g[prefix + '_indptr']), shape=tuple(g[prefix + '_shape']))


def same_csr(a, b):
    assert a.shape == b.shape and a.indices.dtype == b.indices.dtype
    assert np.array_equal(a.indptr, b.indptr)
    assert np.array_equal(a.indices, b.indices)
    assert np.array_equal(a.data, b.data)


@pytest.mark.parametrize('case', ['perturbed_s4', 'graded', 'lshape_u3'])
def test_full_case_bit_exact(case):
    g = load(case)
    c, e, u, x = g['coordinates'], g['elements'], g['u'], g['x']
    same_csr(orc.stiffness_coo(c, e).tocsr(), csr_of(g, 'stiff'))
    same_csr(orc.mass_coo(c, e).tocsr(), csr_of(g, 'mass'))
    i, j, d = orc.mass_triplets(c, e)
    assert np.array_equal(i, g['mass_I']) and np.array_equal(j, g['mass_J'])
    assert np.array_equal(d, g['mass_D'])
    for r in ('MIDPOINT', 'DAYTAYLOR'):
        w, p = g['rule_%s_w' % r], g['rule_%s_p' % r]
        same_csr(orc.general_stiffness_coo(c, e, H.a11, H.a12, H.a21, H.a22,
                                           w, p).tocsr(), csr_of(g, 'gen_' + r))
        same_csr(orc.weighted_mass_coo(c, e, u, H.phi, w, p).tocsr(),
                 csr_of(g, 'wmass_' + r))
        assert np.array_equal(orc.load_vector_cubature(c, e, H.f_load, w, p),
                              g['rhs_' + r])
        assert np.array_equal(orc.load_vector_composition(H.f_nl, u, c, e, w, p),
                              g['nl_load_' + r])
        assert orc.integrate_composition(H.f_nl, u, c, e, w, p) == float(g['nl_integral_' + r])
    assert np.array_equal(orc.load_vector_midpoint(c, e, H.f_load),
                          g['rhs_legacy'])
    e2e, e2n, b2e = orc.geometric_data(e, [g['dirichlet'], g['neumann']])
    assert np.array_equal(e2e, g['element2edges'])
    assert np.array_equal(e2n, g['edge2nodes'])
    assert np.array_equal(b2e[0], g['dir2edges'])
    assert np.array_equal(b2e[1], g['neu2edges'])
    assert np.array_equal(
        orc.eta_r(x, c, e, g['dirichlet'], g['neumann'], H.f_load, H.g_neu),
        g['eta'])
    allb = np.vstack([g['dirichlet'], g['neumann']])
    assert np.array_equal(
        orc.eta_r(x, c, e, allb, np.zeros((0, 2), dtype=int), H.f_load,
                  H.g_neu), g['eta_dirichlet_only'])
    assert np.array_equal(orc.signed_area(c, e), g['area'])


def test_laplace_example_matlab_indicators():
    """test_indicators.py:12-55: eta_R against MATLAB's indicators.dat."""
    g = load('laplace_example')
    eta = orc.eta_r(g['x_reference'], g['coordinates'], g['elements'],
                    g['dirichlet'], g['neumann'], L.f, L.g)
    assert np.array_equal(eta, g['eta_reference'])
    assert np.allclose(eta, g['eta_matlab'])
    same_csr(orc.stiffness_coo(g['coordinates'], g['elements']).tocsr(),
             csr_of(g, 'stiff'))
    # test_solver.py:53-57: energy = x.A.x against MATLAB
    a = orc.stiffness_coo(g['coordinates3'], g['elements3']).tocsr()
    xm = g['x_matlab']
    assert np.allclose(g['x3_reference'], xm)
    assert np.isclose(xm.dot(a.dot(xm)), g['energy_matlab'])
    assert np.isclose(float(g['energy3_reference']), g['energy_matlab'])
    assert np.array_equal(
        orc.load_vector_midpoint(g['coordinates'], g['elements'], L.f),
        g['rhs_legacy'])


def test_ahw_mass_triplets_matlab():
    """test_solver.py:59-86 (triplet ORDER pinned), test_mesh.py:114-130."""
    g = load('ahw_example')
    i, j, d = orc.mass_triplets(g['coordinates'], g['elements'])
    assert np.array_equal(i, g['mass_I_matlab'])
    assert np.array_equal(j, g['mass_J_matlab'])
    assert np.allclose(d, g['mass_D_matlab'])
    assert np.array_equal(d, g['mass_D_reference'])
    assert np.allclose(orc.signed_area(g['coordinates'], g['elements']),
                       g['area_matlab'])


def test_edge_numbering_known_answers():
    """test_mesh.py:29-112."""
    g = load('edge_numbering')
    for pre, nb in (('sq', 2), ('l', 3)):
        bnds = [g['%s_b%d' % (pre, k)] for k in range(nb)]
        e2e, e2n, b2e = orc.geometric_data(g[pre + '_elements'], bnds)
        assert np.array_equal(e2e, g[pre + '_element2edges'])
        assert np.array_equal(e2n, g[pre + '_edge2nodes'])
        for k in range(nb):
            assert np.array_equal(b2e[k], g['%s_b2e%d' % (pre, k)])


def test_sanity_energy_exactly_four():
    """test_sanity_check.py:80-119: assertEqual(4.0, x.A.x)."""
    g = load('sanity_check')
    for level in range(4):
        c, e = g['c%d' % level], g['e%d' % level]
        x = L.sanity_u(c)
        a = orc.stiffness_coo(c, e)
        assert x.dot(a.dot(x)) == 4.0 == float(g['energy%d' % level])


def test_general_stiffness_minus_identity_is_stiffness():
    """test_general_stiffness_matrices.py:22-59."""
    c, e, _ = H.perturbed_square(3)
    w, p = np.array([0.5]), np.array([[1. / 3., 1. / 3.]])
    minus_one = lambda r: -np.ones(r.shape[0])   # noqa: E731
    zero = lambda r: np.zeros(r.shape[0])        # noqa: E731
    gen = orc.general_stiffness_coo(c, e, minus_one, zero, zero, minus_one,
                                    w, p).toarray()
    assert np.allclose(gen, orc.stiffness_coo(c, e).toarray())
