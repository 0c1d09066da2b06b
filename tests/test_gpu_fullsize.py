"""BASELINE.json's full size (S(12): 64M triangles, 33.6M nodes) on the GPU,
checked through size-independent properties -- the oracle needs ~30 GB and
minutes at this size, so it is only used on a window of rows:
  * nnz, N, M match the closed forms of the uniform NVB family;
  * every row sorted and duplicate-free, diagonal present;
  * stiffness: row sums vanish, matrix symmetric (A x . y == x . A y), constants
    in the kernel, energy of a P1-exact function = |grad|^2 * area;
  * mass: entries sum to the area of the domain;
  * a window of rows equals the oracle restricted to those rows bit for bit in
    the indices and to 1e-12 in the values;
  * load vectors of f = 1 add up to the area; eta_R of a P1-exact function equals
    |T|^2 exactly; the edge numbering satisfies Euler's relation, uses every id,
    interior edges twice and boundary edges once.
"""
import numpy as np
import pytest
import torch

from oracle import p1_oracle as orc
from p1afempy_b200 import meshgen

pytestmark = pytest.mark.gpu
LEVEL = 12


@pytest.fixture(scope='module')
def big():
    from p1afempy_b200.device import DeviceMesh
    c, e, boundaries = meshgen.unit_square(LEVEL)
    mesh = DeviceMesh(c, e, reuse=True)
    mesh.test_boundaries = boundaries
    yield c, e, mesh
    mesh.close()


def _spmv(indptr, indices, data, x):
    rows = torch.repeat_interleave(torch.arange(indptr.numel() - 1, device=x.device),
                                   (indptr[1:] - indptr[:-1]).long())
    out = torch.zeros_like(x)
    out.index_add_(0, rows, data * x[indices.long()])
    return out


def test_stiffness_and_mass_at_64m(big):
    c, e, mesh = big
    n = 2 ** LEVEL
    assert e.shape[0] == 4 * n * n and c.shape[0] == 2 * n * n + 2 * n + 1
    ip, ix, a = mesh.stiffness_csr()
    assert mesh.nnz == 14 * n * n + 6 * n + 1 == a.numel()
    lens = (ip[1:] - ip[:-1]).long()
    assert int(lens.min()) >= 4 and int(lens.max()) <= 9
    # sorted, duplicate-free columns inside every row
    inner = torch.ones(ix.numel(), dtype=torch.bool, device=ix.device)
    inner[ip[:-1].long()] = False
    assert bool((ix[1:] > ix[:-1])[inner[1:]].all())
    dev = a.device
    ones = torch.ones(c.shape[0], dtype=torch.float64, device=dev)
    assert float(_spmv(ip, ix, a, ones).abs().max()) < 1e-8          # constants in the kernel
    g = torch.Generator(device='cpu').manual_seed(0)
    x = torch.rand(c.shape[0], generator=g, dtype=torch.float64).to(dev)
    y = torch.rand(c.shape[0], generator=g, dtype=torch.float64).to(dev)
    lhs, rhs = torch.dot(_spmv(ip, ix, a, x), y), torch.dot(x, _spmv(ip, ix, a, y))
    assert abs(float(lhs - rhs)) <= 1e-9 * abs(float(lhs))            # symmetry
    u = 3. * mesh.coords[:, 0] - 2. * mesh.coords[:, 1]              # |grad u|^2 = 13, area 1
    assert abs(float(torch.dot(u, _spmv(ip, ix, a, u))) - 13.) < 1e-9
    # a window of rows against the oracle (needs only the elements touching them)
    lo, hi = 1_000_000, 1_002_000
    touch = np.flatnonzero(((e >= lo) & (e < hi)).any(axis=1))
    ref = orc.stiffness_coo(c, e[touch]).tocsr()[lo:hi]
    s, t = int(ip[lo]), int(ip[hi])
    assert np.array_equal((ip[lo:hi + 1] - ip[lo]).cpu().numpy(), ref.indptr)
    assert np.array_equal(ix[s:t].cpu().numpy(), ref.indices)
    assert np.abs(a[s:t].cpu().numpy() - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
    del a
    _, _, m = mesh.mass_csr(pattern=False)                            # re-assembly: values only
    assert abs(float(m.sum()) - 1.0) < 1e-12
    assert float(m.min()) > 0.


def test_vectors_edges_and_estimator_at_64m(big):
    """The cold routes of the non-matrix operators at full size: plan-free load
    vectors (payload slots), topology-only plan with the slot of every half-edge,
    twins by gather, edge numbering -- checked by exact invariants."""
    from p1afempy_b200 import cubature
    from p1afempy_b200.device import DeviceMesh
    c, e, kept = big
    kept.close()                                        # make room: 64M-sized scratch below
    n = 2 ** LEVEL
    m_el, n_nodes = e.shape[0], c.shape[0]
    with DeviceMesh(kept.coords, kept.elements, check=False) as mesh:
        ones = torch.ones(m_el, dtype=torch.float64, device='cuda')
        # load vectors of f = 1: the entries add up to the area of the domain
        b = mesh.load_vector_midpoint(ones)
        assert mesh._plan is None and abs(float(b.sum()) - 1.0) < 1e-12 and float(b.min()) > 0.
        w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.SMPLX1)
        fq = torch.ones((len(w), m_el), dtype=torch.float64, device='cuda')
        bq = mesh.load_vector(fq, w, p)
        assert mesh._plan is None and abs(float(bq.sum()) - 1.0) < 1e-12
        del fq, bq
        # the estimator of a P1-exact function with f = 1: all jumps vanish exactly on
        # the dyadic mesh, eta_T = (0.5 * 2|T| * 1)^2 = |T|^2 with |T| = 1/M
        x = 3. * mesh.coords[:, 0] - 2. * mesh.coords[:, 1]
        bd = torch.from_numpy(np.vstack(kept.test_boundaries).astype(np.int32)).cuda()
        none = torch.zeros((0, 2), dtype=torch.int32, device='cuda')
        eta = mesh.eta_r(x, bd, none, None, ones)
        assert mesh._plan_topology
        area2 = (1.0 / m_el) ** 2
        assert bool((eta == area2).all())
        del eta
        # edge numbering: Euler's relation, every id used, interior edges twice
        e2e, e2n, b2e = mesh.geometric_data([bd])
        n_edges = e2n.shape[0]
        assert n_edges == n_nodes + m_el - 1
        counts = torch.bincount(e2e.reshape(-1), minlength=n_edges)
        assert int(counts.min()) == 1 and int(counts.max()) == 2
        assert int((counts == 1).sum()) == bd.shape[0] == 4 * n
        assert bool((e2n[:, 0] < e2n[:, 1]).all())
        assert bool((counts[b2e[0]] == 1).all())
