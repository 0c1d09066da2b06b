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


# --------------------------------------------------------------------------
# Row / element windows against the oracle at full size, with the NON-TRIVIAL data
# of configurations C2 / C3: the oracle runs on the
# elements that touch the window (all the window's rows need), the GPU on the whole
# mesh.  Index arrays bit for bit, values to 1e-12 of the row maximum.
# --------------------------------------------------------------------------
import helpers as H  # noqa: E402


def _window_rows(e, lo, hi):
    return np.flatnonzero(((e >= lo) & (e < hi)).any(axis=1))


def _check_window(ip, ix, data, ref, lo, hi, rtol=1e-12):
    s, t = int(ip[lo]), int(ip[hi])
    piece = ref[lo:hi]
    assert np.array_equal((ip[lo:hi + 1] - ip[lo]).cpu().numpy(), piece.indptr)
    assert np.array_equal(ix[s:t].cpu().numpy(), piece.indices)
    H.values_close_by_row(data[s:t].cpu().numpy(), piece, rtol)


WINDOWS = ((0, 1500), (1_000_000, 1_001_500), (33_000_000, 33_001_500))


def test_general_stiffness_and_weighted_mass_windows_at_64m(big):
    """C3: a11 = -(1+x^2), a12 = -xy/4, a21 = -xy/2, a22 = -(2+sin(pi y)); weighted
    mass with u = sin(pi x) sin(pi y), phi(u) = 1 + u^2; 16-point rule; cold path."""
    from p1afempy_b200 import cubature
    from p1afempy_b200.device import DeviceMesh
    c, e, kept = big
    kept.close()
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.DAYTAYLOR)
    pi = float(np.pi)
    with DeviceMesh(kept.coords, kept.elements, check=False) as mesh:
        xq = mesh.quadrature_points(p)
        qx, qy = xq[..., 0], xq[..., 1]
        qs = [(-(1. + qx ** 2)).contiguous(), (-0.25 * qx * qy).contiguous(),
              (-0.5 * qx * qy).contiguous(), (-(2. + torch.sin(pi * qy))).contiguous()]
        del xq, qx, qy
        ip, ix, g = mesh.general_stiffness_csr(*qs, w)
        assert g.numel() == 14 * 4 ** LEVEL + 6 * 2 ** LEVEL + 1
        for lo, hi in WINDOWS:
            touch = _window_rows(e, lo, hi)
            ref = orc.general_stiffness_coo(c, e[touch], H.a11, H.a12, H.a21, H.a22, w, p).tocsr()
            _check_window(ip, ix, g, ref, lo, hi)
        del qs, g, ip, ix
        u_host = H.u_nodal(c)
        u = torch.from_numpy(u_host).cuda()
        uq = mesh.nodal_at_quadrature(u, p)
        phiq = (1. + uq ** 2).contiguous()
        del uq
        ip, ix, wm = mesh.weighted_mass_csr(phiq, w, p)
        for lo, hi in WINDOWS:
            touch = _window_rows(e, lo, hi)
            ref = orc.weighted_mass_coo(c, e[touch], u_host, H.phi, w, p).tocsr()
            _check_window(ip, ix, wm, ref, lo, hi)


def test_load_vector_and_eta_windows_at_64m(big):
    """Cubature load vector of f = sin(2 pi x) cos(3 pi y) + 1 (values from CUDA's
    libm: 1e-12) and of a polynomial f (bit for bit), and eta_R of u = sin sin with
    f at the centroids, on windows of nodes / elements."""
    from p1afempy_b200 import cubature
    from p1afempy_b200.device import DeviceMesh
    c, e, kept = big
    kept.close()
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.DAYTAYLOR)
    pi = float(np.pi)
    poly = lambda r: 2. + r[:, 0] * r[:, 1] - 0.5 * r[:, 1] ** 2      # noqa: E731
    with DeviceMesh(kept.coords, kept.elements, check=False) as mesh:
        xq = mesh.quadrature_points(p)
        qx, qy = xq[..., 0], xq[..., 1]
        fq = (torch.sin(2. * pi * qx) * torch.cos(3. * pi * qy) + 1.).contiguous()
        fq_poly = (2. + qx * qy - 0.5 * qy ** 2).contiguous()
        del xq, qx, qy
        b = mesh.load_vector(fq, w, p)
        b_poly = mesh.load_vector(fq_poly, w, p)
        del fq, fq_poly
        for lo, hi in WINDOWS:
            touch = _window_rows(e, lo, hi)
            ref = orc.load_vector_cubature(c, e[touch], H.f_load, w, p)[lo:hi]
            got = b[lo:hi].cpu().numpy()
            assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
            ref = orc.load_vector_cubature(c, e[touch], poly, w, p)[lo:hi]
            assert np.array_equal(b_poly[lo:hi].cpu().numpy(), ref)
        del b, b_poly
        # eta_R: x and f(centroid) are arrays -> fed bit-identically; windows of elements
        x_host = H.u_nodal(c) + 0.25 * c[:, 0]
        x = torch.from_numpy(x_host).cuda()
        cen = mesh.centroids()
        fc = (2. + cen[:, 0] * cen[:, 1] - 0.5 * cen[:, 1] ** 2).contiguous()
        bd = torch.from_numpy(np.vstack(kept.test_boundaries).astype(np.int32)).cuda()
        none = torch.zeros((0, 2), dtype=torch.int32, device='cuda')
        eta = mesh.eta_r(x, bd, none, None, fc)
        for t0 in (0, 20_000_000, 67_000_000):
            t1 = t0 + 1000
            nodes = np.unique(e[t0:t1])
            mask = np.zeros(c.shape[0], dtype=bool)
            mask[nodes] = True
            sub = np.flatnonzero(mask[e].any(axis=1))        # the window and its ring
            es = e[sub]
            # every boundary edge of the sub-mesh is listed (artificial ones included:
            # their jump is zeroed, which only touches ring elements)
            he = np.concatenate([es[:, [0, 1]], es[:, [1, 2]], es[:, [2, 0]]])
            key = np.sort(he, axis=1)
            _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
            bsub = he[cnt[np.asarray(inv).reshape(-1)] == 1]
            ref = orc.eta_r(x_host, c, es, bsub, np.zeros((0, 2), dtype=np.int64), poly,
                            lambda r: 0. * r[:, 0])
            pos = np.searchsorted(sub, np.arange(t0, t1))
            assert np.array_equal(sub[pos], np.arange(t0, t1))
            assert np.array_equal(eta[t0:t1].cpu().numpy(), ref[pos])


def test_mass_and_load_vector_windows_at_16m():
    """C2 (S(11), 16.8M triangles): mass matrix to CSR and the cubature load vector."""
    from p1afempy_b200 import cubature
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = meshgen.unit_square(11)
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.DAYTAYLOR)
    pi = float(np.pi)
    with DeviceMesh(c, e) as mesh:
        ip, ix, m = mesh.mass_csr()
        xq = mesh.quadrature_points(p)
        fq = (torch.sin(2. * pi * xq[..., 0]) * torch.cos(3. * pi * xq[..., 1]) + 1.).contiguous()
        del xq
        b = mesh.load_vector(fq, w, p)
        for lo, hi in ((0, 2000), (4_000_000, 4_002_000), (8_390_000, 8_392_000)):
            touch = _window_rows(e, lo, hi)
            _check_window(ip, ix, m, orc.mass_coo(c, e[touch]).tocsr(), lo, hi)
            ref = orc.load_vector_cubature(c, e[touch], H.f_load, w, p)[lo:hi]
            assert np.abs(b[lo:hi].cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(float(m.sum()) - 1.0) < 1e-12


def test_stiffness_window_and_int64_indices_at_256m():
    """The top of the sweep, S(13): 268M triangles, nnz = 939.6M -- scipy's rule
    (max(9 M, N) > 2^31 - 1, scipy/sparse/_coo.py:419) makes the reference's CSR
    index arrays int64 there.  The device assembles with int32 indices and widens
    them (p1_widen_indices); windows of rows against the oracle."""
    import ctypes as C
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import DeviceMesh, scipy_index_dtype, _ptr, _stream
    free, total = torch.cuda.mem_get_info()
    if total < 120 << 30:
        pytest.skip('needs a 180 GB device')
    level = 13
    n = 2 ** level
    ct, et, _ = meshgen.unit_square_device(level)
    m_el, n_nodes = et.shape[0], ct.shape[0]
    assert m_el == 4 * n * n and n_nodes == 2 * n * n + 2 * n + 1
    assert scipy_index_dtype(m_el, n_nodes, n_nodes) == np.int64
    with DeviceMesh(ct, et, check=False) as mesh:
        ip, ix, a = mesh.stiffness_csr()
        assert mesh.nnz == 14 * n * n + 6 * n + 1 == a.numel()
        ix64 = torch.empty(ix.shape, dtype=torch.int64, device='cuda')
        ip64 = torch.empty(ip.shape, dtype=torch.int64, device='cuda')
        _lib.check(mesh._lib.p1_widen_indices(mesh.ctx, _ptr(ix), ix.numel(), _ptr(ix64),
                                              _stream(mesh.device)))
        _lib.check(mesh._lib.p1_widen_indices(mesh.ctx, _ptr(ip), ip.numel(), _ptr(ip64),
                                              _stream(mesh.device)))
        assert int(ip64[-1]) == mesh.nnz and int(ix64.max()) == n_nodes - 1
        # windows: the elements that touch them, found on the device
        c_host = None
        for lo, hi in ((0, 1000), (50_000_000, 50_001_000), (134_233_000, 134_234_000)):
            touch = torch.nonzero(((et >= lo) & (et < hi)).any(dim=1)).reshape(-1)
            es = et[touch].cpu().numpy().astype(np.int64)
            if c_host is None:
                c_host = ct.cpu().numpy()
            ref = orc.stiffness_coo(c_host, es).tocsr()[lo:hi]
            s, t = int(ip64[lo]), int(ip64[hi])
            assert np.array_equal((ip64[lo:hi + 1] - ip64[lo]).cpu().numpy(), ref.indptr)
            got_ix = ix64[s:t].cpu().numpy()
            assert got_ix.dtype == np.int64 and np.array_equal(got_ix, ref.indices)
            H.values_close_by_row(a[s:t].cpu().numpy(), ref, 1e-12)
        del ix64, ip64
        ones = torch.ones(n_nodes, dtype=torch.float64, device='cuda')
        rows = torch.repeat_interleave(torch.arange(n_nodes, device='cuda'),
                                       (ip[1:] - ip[:-1]).long())
        out = torch.zeros_like(ones)
        out.index_add_(0, rows, a)
        assert float(out.abs().max()) < 1e-7            # row sums vanish
