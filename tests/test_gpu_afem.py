"""The callers either side of the assembly path on the GPU:
refineNVB, get_element_to_neighbours, the device side of solve_laplace, Doerfler
marking -- against fixtures produced by the REAL reference
(tests/golden/make_golden.py: afem_case, laplace_case), which carry the reference's
own golden files (tests/data/refined_nvb/*, get_neighbours/*, laplace_example/*)."""
import os

import numpy as np
import pytest
import torch

from oracle import p1_oracle as orc
from p1afempy_b200 import marking, meshgen
import helpers as H
import helpers_laplace as L

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(H.GOLDEN, 'afem.npz'))


def test_refine_nvb_reference_cases():
    """test_refinement.py:11-114: the 2-triangle square with both elements marked
    (hard-coded answer) and the L-shape with elements 0, 1, 3, 5 marked (the
    reference's fixture files)."""
    from p1afempy_b200 import refinement
    c = np.array([[0., 0.], [1., 0.], [1., 1.], [0., 1.]])
    e = np.array([[0, 1, 2], [0, 2, 3]])
    bcs = [np.array([[0, 1], [1, 2]]), np.array([[2, 3], [3, 0]])]
    rc, re_, rb, emb = refinement.refineNVB(coordinates=c, elements=e,
                                            marked_elements=np.array([0, 1]),
                                            boundary_conditions=bcs)
    assert np.array_equal(rc, [[0., 0.], [1., 0.], [1., 1.], [0., 1.], [0.5, 0.], [0.5, 0.5],
                               [1., 0.5], [0.5, 1.], [0., 0.5]])
    assert np.array_equal(re_, [[4, 2, 5], [0, 4, 5], [4, 1, 6], [2, 4, 6], [5, 3, 8], [0, 5, 8],
                                [5, 2, 7], [3, 5, 7]])
    assert np.array_equal(rb[0], [[0, 4], [1, 6], [4, 1], [6, 2]])
    assert np.array_equal(rb[1], [[2, 7], [3, 8], [7, 3], [8, 0]])
    assert emb.size == 0 and re_.dtype == np.int64
    rc, re_, rb, _ = refinement.refineNVB(
        coordinates=G['l_c'], elements=G['l_e'], marked_elements=G['l_marked'],
        boundary_conditions=[G['l_b0'], G['l_b1'], G['l_b2']])
    assert np.array_equal(rc, G['l_rc']) and np.array_equal(rc, G['l_rc_file'])
    assert np.array_equal(re_, G['l_re']) and np.array_equal(re_, G['l_re_file'])
    for j in range(3):
        assert np.array_equal(rb[j], G['l_rb%d' % j])


@pytest.mark.parametrize('tag', ['plain', 'sorted'])
def test_refine_nvb_adaptive_steps_with_embedding(tag):
    """Four adaptive steps (30 % of the elements marked at random) with a nodal
    vector embedded, in both element layouts (sort_for_longest_egde rotates the
    caller's elements in place, refinement.py:556-567): node numbers, element
    order, boundaries and embedded values are the reference's bit for bit."""
    from p1afempy_b200 import refinement
    nb = int(G[tag + '_nb'])
    for step in range(4):
        e_in = G['%s_in_e%d' % (tag, step)].copy()
        bcs = [G['%s_in_b%d_%d' % (tag, step, j)] for j in range(nb)]
        c, e, b, emb = refinement.refineNVB(
            coordinates=G['%s_in_c%d' % (tag, step)], elements=e_in,
            marked_elements=G['%s_marked%d' % (tag, step)], boundary_conditions=bcs,
            to_embed=G['%s_in_emb%d' % (tag, step)], sort_for_longest_egde=(tag == 'sorted'))
        assert np.array_equal(c, G['%s_c%d' % (tag, step)])
        assert np.array_equal(e, G['%s_e%d' % (tag, step)])
        assert np.array_equal(emb, G['%s_emb%d' % (tag, step)])
        for j in range(nb):
            assert np.array_equal(b[j], G['%s_b%d_%d' % (tag, step, j)])
        if tag == 'sorted':   # longest edge first, in place
            cc = G['%s_in_c%d' % (tag, step)]
            d = cc[e_in[:, [1, 2, 0]]] - cc[e_in]
            l2 = (d ** 2).sum(axis=2)
            assert np.all(l2[:, 0] >= l2[:, 1]) and np.all(l2[:, 0] >= l2[:, 2])
        else:
            assert np.array_equal(e_in, G['%s_in_e%d' % (tag, step)])


def test_element_to_neighbours():
    """mesh.py:495-528: the reference's MATLAB fixture (test_mesh.py:504-530) and a
    graded mesh."""
    from p1afempy_b200 import mesh
    nb = mesh.get_element_to_neighbours(G['nb_elements_matlab'])
    assert nb.dtype == np.int64
    assert np.array_equal(nb, G['nb_matlab']) and np.array_equal(nb, G['nb_reference_matlab'])
    assert np.array_equal(mesh.get_element_to_neighbours(G['nb_elements_graded']),
                          G['nb_reference_graded'])


def test_solve_laplace_on_the_device():
    """test_solver.py:20-57: MATLAB's x and energy.  solver='direct' is the
    reference's route (scipy's direct solver on the free nodes) with assembly, load
    vector, Dirichlet lift and Neumann load on the device; solver='pcg' keeps the
    matrix on the GPU."""
    from p1afempy_b200 import solvers
    g = np.load(os.path.join(H.GOLDEN, 'laplace_example.npz'))
    args = dict(coordinates=g['coordinates3'], elements=g['elements3'],
                dirichlet=g['dirichlet3'], neumann=g['neumann3'], f=L.f, g=L.g, uD=L.uD)
    x, energy = solvers.solve_laplace(**args)
    assert np.allclose(x, g['x_matlab']) and np.isclose(energy, float(g['energy_matlab']))
    assert np.abs(x - g['x3_reference']).max() <= 1e-12 * np.abs(g['x3_reference']).max()
    assert abs(energy - float(g['energy3_reference'])) <= 1e-12 * abs(float(g['energy3_reference']))
    xp, ep, info = solvers.solve_laplace(**args, solver='pcg', tol=1e-13, return_info=True)
    assert info['iterations'] > 0 and info['relative_residual'] <= 1e-12
    assert np.allclose(xp, g['x_matlab']) and np.isclose(ep, float(g['energy_matlab']))
    assert np.abs(xp - g['x3_reference']).max() <= 1e-10 * np.abs(g['x3_reference']).max()
    # a larger problem: the two routes agree
    c, e, b = H.perturbed_square(5)
    allb = np.vstack(b)
    args = dict(coordinates=c, elements=e, dirichlet=allb[:70], neumann=allb[70:],
                f=H.f_load, g=H.g_neu, uD=lambda r: r[:, 0] - 2. * r[:, 1])
    xd, ed = solvers.solve_laplace(**args)
    xp, ep = solvers.solve_laplace(**args, solver='pcg', tol=1e-13)
    assert np.abs(xd - xp).max() <= 1e-9 * np.abs(xd).max() and abs(ed - ep) <= 1e-9 * abs(ed)


def test_apply_neumann_matches_reference_arithmetic():
    """solvers.py:601-618: |E| g(mid) / 2 added at both endpoints in bincount order."""
    from p1afempy_b200 import solvers
    c, e, b = H.perturbed_square(4)
    neu = np.vstack(b)[5:40]
    rhs = np.linspace(-1., 2., c.shape[0])
    got = solvers.apply_neumann(neumann_bc=neu, coordinates=c, g=H.g_neu, b=rhs)
    cn1, cn2 = c[neu[:, 0]], c[neu[:, 1]]
    ref = rhs + np.bincount(neu.flatten(order='F'),
                            weights=np.tile(np.sqrt(np.sum(np.square(cn2 - cn1), axis=1))
                                            * H.g_neu((cn1 + cn2) / 2) / 2., (2, 1)).flatten(),
                            minlength=rhs.size)
    assert np.array_equal(got, ref)


def test_doerfler_marking():
    rng = np.random.default_rng(1)
    eta = rng.random(5000) ** 4
    eta[100:110] = eta[100]                 # ties: by index
    for theta in (0.1, 0.5, 0.9, 1.0):
        got = marking.doerfler_marking(eta, theta)
        order = np.argsort(-eta, kind='stable')
        csum = np.cumsum(eta[order])
        k = int(np.searchsorted(csum, theta * csum[-1])) + 1
        # the device's prefix sum is a parallel scan: where the partial sums pass the
        # threshold within rounding, the count may differ from numpy's sequential one
        n = min(len(got), k)
        assert np.array_equal(got[:n], order[:n])                 # same order, ties by index
        assert csum[len(got) - 1] >= theta * csum[-1] * (1 - 1e-12)          # enough ...
        assert len(got) == 1 or csum[len(got) - 2] <= theta * csum[-1] * (1 + 1e-12)  # ... and minimal
    assert marking.doerfler_marking(np.zeros(0)).size == 0


def test_adaptive_loop_on_the_device():
    """solve -- estimate -- mark -- refine, everything device-resident: the meshes
    equal the host generator's (== the reference's refineNVB)."""
    from p1afempy_b200.device import DeviceMesh
    from p1afempy_b200.refinement import refine_nvb_device
    c, e, b = meshgen.l_shape()
    c, e, b = meshgen.refine_uniform(c, e, b, times=3)
    ct = torch.from_numpy(c).cuda()
    et = torch.from_numpy(e.astype(np.int32)).cuda()
    bt = [torch.from_numpy(x.astype(np.int32)).cuda() for x in b]
    none = torch.zeros((0, 2), dtype=torch.int32, device='cuda')
    for _ in range(4):
        x = (ct[:, 0] - 1.) ** 2 * ct[:, 1]
        with DeviceMesh(ct, et, check=False) as mesh:
            eta = mesh.eta_r(x, torch.cat(bt), none, None,
                             torch.ones(et.shape[0], dtype=torch.float64, device='cuda'))
        marked = marking.doerfler_marking_device(eta, 0.4)
        c, e, b = meshgen.refine_nvb(c, e, marked.cpu().numpy(), b)
        ct, et, bt, _ = refine_nvb_device(ct, et, marked, bt)
        assert np.array_equal(ct.cpu().numpy(), c) and np.array_equal(et.cpu().numpy(), e)
        for u, v in zip(bt, b):
            assert np.array_equal(u.cpu().numpy(), v)
