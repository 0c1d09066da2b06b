"""meshgen.refine_nvb / number_edges must reproduce the reference's refineNVB
and provide_geometric_data bit for bit (only runs where /root/reference is
mounted; on the GPU box these are skipped)."""
import numpy as np
import pytest

from oracle import ref_import
from p1afempy_b200 import meshgen

pytestmark = pytest.mark.skipif(not ref_import.available(),
                                reason='reference not mounted')


def test_uniform_square_matches_refineNVB():
    ref = ref_import.load()
    c, e, b = meshgen.criss_square()
    rc, re_, rb = c.copy(), e.copy(), [x.copy() for x in b]
    for _ in range(5):
        c, e, b = meshgen.refine_nvb(c, e, np.arange(e.shape[0]), b)
        rc, re_, rb, _ = ref.refinement.refineNVB(
            rc, re_, np.arange(re_.shape[0]), rb)
        assert np.array_equal(c, rc)
        assert np.array_equal(e, re_)
        assert all(np.array_equal(x, y) for x, y in zip(b, rb))
    c2, e2, b2 = meshgen.unit_square(5)
    assert np.array_equal(c2, rc) and np.array_equal(e2, re_)
    n = 2 ** 5
    assert e2.shape[0] == 4 * n * n and c2.shape[0] == 2 * n * n + 2 * n + 1


def test_adaptive_lshape_matches_refineNVB():
    ref = ref_import.load()
    c, e, b = meshgen.l_shape()
    rc, re_, rb = c.copy(), e.copy(), [x.copy() for x in b]
    rng = np.random.default_rng(1)
    for _ in range(8):
        marked = np.flatnonzero(rng.random(e.shape[0]) < 0.3)
        c, e, b = meshgen.refine_nvb(c, e, marked, b)
        rc, re_, rb, _ = ref.refinement.refineNVB(rc, re_, marked, rb)
        assert np.array_equal(c, rc)
        assert np.array_equal(e, re_)
        assert all(np.array_equal(x, y) for x, y in zip(b, rb))
    e2e, e2n, b2e = meshgen.number_edges(e, b)
    r2e, r2n, rb2e = ref.mesh.provide_geometric_data(re_, rb)
    assert np.array_equal(e2e, r2e) and np.array_equal(e2n, r2n)
    assert all(np.array_equal(x, y) for x, y in zip(b2e, rb2e))


def test_lshape_fixture_identical():
    ref = ref_import.load()
    from pathlib import Path
    d = Path(ref_import.data_dir()) / 'l_shape_mesh'
    rc, re_ = ref.io_helpers.read_mesh(
        path_to_coordinates=d / 'l_shape_coordinates.dat',
        path_to_elements=d / 'l_shape_elements.dat')
    c, e, b = meshgen.l_shape()
    assert np.array_equal(c, rc) and np.array_equal(e, re_)
    for k in range(3):
        rb = ref.io_helpers.read_boundary_condition(d / f'l_shape_bc_{k}.dat')
        assert np.array_equal(b[k], rb)
