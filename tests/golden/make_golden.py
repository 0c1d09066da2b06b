"""Generates tests/golden/*.npz by running the REAL reference
(/root/reference, leuraph/p1afempy v0.2.16) in the build container.

    python tests/golden/make_golden.py

Each file stores the inputs and the reference's outputs for one case, plus --
where the reference's own tests hold them -- the MATLAB golden vectors
(laplace_example/{x,energy,indicators}.dat, ahw_codes_example_mesh/
{mass_matrix_I,J,D,area}.dat) and the hard-coded known answers of
test_mesh.py:51-111.  The fixtures travel to the GPU box; the reference does
not.  Callbacks are the ones in tests/helpers.py and
/root/reference/p1afempy/tests/auto/example_setup.py (restated in
tests/helpers_laplace.py).
"""
import os
import sys
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

from oracle import ref_import  # noqa: E402
from p1afempy_b200 import cubature  # noqa: E402
import helpers as H  # noqa: E402
import helpers_laplace as L  # noqa: E402

ref = ref_import.load()
DATA = Path(ref_import.data_dir())
io = ref.io_helpers


def csr_parts(prefix, coo):
    a = coo.tocsr()
    return {prefix + '_indptr': a.indptr, prefix + '_indices': a.indices,
            prefix + '_data': a.data,
            prefix + '_shape': np.array(a.shape, dtype=np.int64)}


def full_case(name, c, e, bnds):
    """All hot-path outputs of the reference on one mesh."""
    allb = np.vstack(bnds)
    cut = (3 * allb.shape[0]) // 5
    dirichlet, neumann = allb[:cut], allb[cut:]
    u = H.u_nodal(c)
    x = u + 0.1 * c[:, 0]
    out = dict(coordinates=c, elements=e, dirichlet=dirichlet,
               neumann=neumann, u=u, x=x)
    out.update(csr_parts('stiff', ref.solvers.get_stiffness_matrix(c, e)))
    out.update(csr_parts('mass', ref.solvers.get_mass_matrix(c, e)))
    i, j, d = ref.solvers.get_mass_matrix_elements(c, e)
    out.update(mass_I=i, mass_J=j, mass_D=d)
    for rname in ('MIDPOINT', 'DAYTAYLOR'):
        rule = cubature.CubatureRuleEnum[rname]
        w, p = cubature.resolve_rule(rule)
        out['rule_%s_w' % rname] = w
        out['rule_%s_p' % rname] = p
        out.update(csr_parts('gen_' + rname,
                             ref.solvers.get_general_stiffness_matrix(
                                 c, e, H.a11, H.a12, H.a21, H.a22, rule)))
        out.update(csr_parts('wmass_' + rname,
                             ref.solvers.get_weighted_mass_matrix(
                                 c, e, u, H.phi, rule)))
        out['rhs_' + rname] = ref.solvers.get_right_hand_side(
            c, e, H.f_load, cubature_rule=rule)
        out['nl_load_' + rname] = \
            ref.solvers.get_load_vector_of_composition_nonlinear_with_fem(
                f=H.f_nl, u=u, coordinates=c, elements=e, cubature_rule=rule)
        out['nl_integral_' + rname] = np.array(
            ref.solvers.integrate_composition_nonlinear_with_fem(
                f=H.f_nl, u=u, coordinates=c, elements=e, cubature_rule=rule))
    out['rhs_legacy'] = ref.solvers.get_right_hand_side(c, e, H.f_load)
    e2e, e2n, b2e = ref.mesh.provide_geometric_data(e, [dirichlet, neumann])
    out.update(element2edges=e2e, edge2nodes=e2n, dir2edges=b2e[0],
               neu2edges=b2e[1])
    out['eta'] = ref.indicators.compute_eta_r(x, c, e, dirichlet, neumann,
                                              H.f_load, H.g_neu)
    out['eta_dirichlet_only'] = ref.indicators.compute_eta_r(
        x, c, e, allb, np.zeros((0, 2), dtype=int), H.f_load, H.g_neu)
    out['area'] = ref.mesh.get_area(c, e)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'M=%d N=%d' % (e.shape[0], c.shape[0]))


def laplace_case():
    """test_indicators.py:12-55 and test_solver.py:20-57."""
    d = DATA / 'laplace_example'
    c, e = io.read_mesh(path_to_coordinates=d / 'coordinates.dat',
                        path_to_elements=d / 'elements.dat')
    bcs = [io.read_boundary_condition(d / 'dirichlet.dat'),
           io.read_boundary_condition(d / 'neumann.dat')]
    for _ in range(2):
        c, e, bcs, _ = ref.refinement.refineNVB(
            coordinates=c, elements=e, marked_elements=np.arange(e.shape[0]),
            boundary_conditions=bcs)
    x, energy = ref.solvers.solve_laplace(
        coordinates=c, elements=e, dirichlet=bcs[0], neumann=bcs[1],
        f=L.f, g=L.g, uD=L.uD)
    eta = ref.indicators.compute_eta_r(x=x, coordinates=c, elements=e,
                                       dirichlet=bcs[0], neumann=bcs[1],
                                       f=L.f, g=L.g)
    out = dict(coordinates=c, elements=e, dirichlet=bcs[0], neumann=bcs[1],
               x_reference=x, energy_reference=np.array(energy),
               eta_reference=eta,
               x_matlab=np.loadtxt(d / 'x.dat'),
               energy_matlab=np.loadtxt(d / 'energy.dat'),
               eta_matlab=np.loadtxt(d / 'indicators.dat'))
    out.update(csr_parts('stiff', ref.solvers.get_stiffness_matrix(c, e)))
    out['rhs_legacy'] = ref.solvers.get_right_hand_side(c, e, L.f)
    # x.dat / energy.dat belong to THREE refinements (test_solver.py:38-57)
    c3, e3, b3, _ = ref.refinement.refineNVB(
        coordinates=c, elements=e, marked_elements=np.arange(e.shape[0]),
        boundary_conditions=bcs)
    x3, energy3 = ref.solvers.solve_laplace(
        coordinates=c3, elements=e3, dirichlet=b3[0], neumann=b3[1],
        f=L.f, g=L.g, uD=L.uD)
    out.update(coordinates3=c3, elements3=e3, dirichlet3=b3[0],
               neumann3=b3[1], x3_reference=x3,
               energy3_reference=np.array(energy3))
    np.savez_compressed(os.path.join(HERE, 'laplace_example.npz'), **out)
    print('laplace_example', e.shape, eta.shape)


def ahw_case():
    """test_solver.py:59-86 and test_mesh.py:114-130 (MATLAB, 1-based)."""
    d = DATA / 'ahw_codes_example_mesh'
    c, e = io.read_mesh(path_to_coordinates=d / 'coordinates.dat',
                        path_to_elements=d / 'elements.dat',
                        shift_indices=True)
    i, j, dd = ref.solvers.get_mass_matrix_elements(c, e)
    out = dict(coordinates=c, elements=e,
               mass_I_reference=i, mass_J_reference=j, mass_D_reference=dd,
               mass_I_matlab=np.loadtxt(d / 'mass_matrix_I.dat') - 1,
               mass_J_matlab=np.loadtxt(d / 'mass_matrix_J.dat') - 1,
               mass_D_matlab=np.loadtxt(d / 'mass_matrix_D.dat'),
               area_matlab=np.loadtxt(d / 'area.dat'),
               area_reference=ref.mesh.get_area(c, e))
    np.savez_compressed(os.path.join(HERE, 'ahw_example.npz'), **out)
    print('ahw_example', e.shape)


def edge_numbering_case():
    """Known answers hard-coded in test_mesh.py:51-63 and :86-111."""
    d = DATA / 'simple_square_mesh'
    _, sq_e = io.read_mesh(path_to_coordinates=d / 'coordinates.dat',
                           path_to_elements=d / 'elements.dat')
    sq_b = [io.read_boundary_condition(d / 'square_boundary_0.dat'),
            io.read_boundary_condition(d / 'square_boundary_1.dat')]
    d = DATA / 'l_shape_mesh'
    _, l_e = io.read_mesh(path_to_coordinates=d / 'l_shape_coordinates.dat',
                          path_to_elements=d / 'l_shape_elements.dat')
    l_b = [io.read_boundary_condition(d / ('l_shape_bc_%d.dat' % k))
           for k in range(3)]
    out = dict(
        sq_elements=sq_e, sq_b0=sq_b[0], sq_b1=sq_b[1],
        sq_element2edges=np.array([[0, 2, 1], [1, 3, 4]]),
        sq_edge2nodes=np.array([[0, 1], [0, 2], [1, 2], [2, 3], [0, 3]]),
        sq_b2e0=np.array([0, 2]), sq_b2e1=np.array([3, 4]),
        l_elements=l_e, l_b0=l_b[0], l_b1=l_b[1], l_b2=l_b[2],
        l_element2edges=np.array([[0, 5, 1], [1, 6, 12], [2, 7, 5],
                                  [3, 8, 7], [6, 9, 11], [4, 10, 9]]),
        l_edge2nodes=np.array([[0, 1], [0, 4], [1, 2], [2, 3], [4, 5],
                               [1, 4], [4, 7], [2, 4], [3, 4], [4, 6],
                               [5, 6], [6, 7], [0, 7]]),
        l_b2e0=np.array([0, 2, 3]), l_b2e1=np.array([8, 4, 10]),
        l_b2e2=np.array([11, 12]))
    # the constants above must be what the reference really returns
    for pre, e, b in (('sq', sq_e, sq_b), ('l', l_e, l_b)):
        e2e, e2n, b2e = ref.mesh.provide_geometric_data(e, b)
        assert np.array_equal(e2e, out[pre + '_element2edges'])
        assert np.array_equal(e2n, out[pre + '_edge2nodes'])
        for k, x in enumerate(b2e):
            assert np.array_equal(x, out['%s_b2e%d' % (pre, k)])
    np.savez_compressed(os.path.join(HERE, 'edge_numbering.npz'), **out)
    print('edge_numbering ok')


def sanity_case():
    """test_sanity_check.py:80-119: energy x.A.x == 4.0 EXACTLY."""
    d = DATA / 'sanity_check'
    c, e = io.read_mesh(path_to_coordinates=d / 'coordinates.dat',
                        path_to_elements=d / 'elements.dat')
    b = [io.read_boundary_condition(d / 'dirichlet.dat')]
    out = {}
    for level in range(4):
        a = ref.solvers.get_stiffness_matrix(c, e)
        x = L.sanity_u(c)
        out['c%d' % level] = c
        out['e%d' % level] = e
        out['energy%d' % level] = np.array(x.dot(a.dot(x)))
        c, e, b, _ = ref.refinement.refineNVB(c, e, np.arange(e.shape[0]), b)
    np.savez_compressed(os.path.join(HERE, 'sanity_check.npz'), **out)
    print('sanity', [float(out['energy%d' % k]) for k in range(4)])


def afem_case():
    """The callers either side of the path: refineNVB with
    to_embed / sort_for_longest_egde (refinement.py:508-718; test_refinement.py:11-114
    and its fixtures tests/data/refined_nvb/*), get_element_to_neighbours
    (mesh.py:495-528; test_mesh.py:504-530 with the MATLAB fixture)."""
    out = {}
    # the reference's own L-shape case
    d = DATA / 'l_shape_mesh'
    c, e = io.read_mesh(path_to_coordinates=d / 'l_shape_coordinates.dat',
                        path_to_elements=d / 'l_shape_elements.dat')
    bcs = [io.read_boundary_condition(d / ('l_shape_bc_%d.dat' % j)) for j in range(3)]
    marked = np.array([0, 1, 3, 5])
    rc, re_, rb, _ = ref.refinement.refineNVB(coordinates=c, elements=e, marked_elements=marked,
                                             boundary_conditions=bcs)
    r = DATA / 'refined_nvb'
    out.update(l_c=c, l_e=e, l_b0=bcs[0], l_b1=bcs[1], l_b2=bcs[2], l_marked=marked,
               l_rc=rc, l_re=re_, l_rb0=rb[0], l_rb1=rb[1], l_rb2=rb[2],
               l_rc_file=np.loadtxt(r / 'l_shape_coordinates_refined.dat'),
               l_re_file=np.loadtxt(r / 'l_shape_elements_refined.dat').astype(np.int64) - 1)
    # several adaptive steps with an embedded nodal vector, both element layouts
    rng = np.random.default_rng(3)
    for tag, sort in (('plain', False), ('sorted', True)):
        c, e, b = H.perturbed_square(3)
        e = e.copy()
        emb = H.u_nodal(c)
        for step in range(4):
            marked = np.flatnonzero(rng.random(e.shape[0]) < 0.3)
            out['%s_in_c%d' % (tag, step)] = c
            out['%s_in_e%d' % (tag, step)] = e.copy()
            out['%s_in_emb%d' % (tag, step)] = emb
            out['%s_marked%d' % (tag, step)] = marked
            for j, x in enumerate(b):
                out['%s_in_b%d_%d' % (tag, step, j)] = x
            c, e, b, emb = ref.refinement.refineNVB(
                coordinates=c, elements=e, marked_elements=marked, boundary_conditions=b,
                to_embed=emb, sort_for_longest_egde=sort)
            out['%s_sorted_e%d' % (tag, step)] = out['%s_in_e%d' % (tag, step)] if not sort else None
            out['%s_c%d' % (tag, step)] = c
            out['%s_e%d' % (tag, step)] = e.copy()   # (the next step sorts e in place)
            out['%s_emb%d' % (tag, step)] = emb
            for j, x in enumerate(b):
                out['%s_b%d_%d' % (tag, step, j)] = x
            out['%s_nb' % tag] = np.array(len(b))
    out = {k: v for k, v in out.items() if v is not None}
    # neighbours: MATLAB fixture and a graded mesh
    g = DATA / 'get_neighbours'
    em = np.loadtxt(g / 'elements_matlab.dat').astype(np.int64) - 1
    out.update(nb_elements_matlab=em,
               nb_matlab=np.loadtxt(g / 'element2neighbours_matlab.dat').astype(np.int64) - 1,
               nb_reference_matlab=ref.mesh.get_element_to_neighbours(em))
    cg, eg, _ = H.graded_square()
    out.update(nb_elements_graded=eg, nb_reference_graded=ref.mesh.get_element_to_neighbours(eg))
    np.savez_compressed(os.path.join(HERE, 'afem.npz'), **out)
    print('afem', len(out))


if __name__ == '__main__':
    full_case('perturbed_s4', *H.perturbed_square(4))
    full_case('graded', *H.graded_square())
    full_case('lshape_u3', *__import__('p1afempy_b200.meshgen', fromlist=['x'])
              .refine_uniform(*__import__('p1afempy_b200.meshgen',
                                          fromlist=['x']).l_shape(), times=3))
    laplace_case()
    ahw_case()
    edge_numbering_case()
    sanity_case()
    afem_case()
