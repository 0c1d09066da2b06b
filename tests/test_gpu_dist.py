"""Row-range (multi-GPU) assembly on the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import p1_oracle as orc
from p1afempy_b200 import dist as pdist
from p1afempy_b200 import meshgen
import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('world,chunk', [(2, 4), (3, 6), (8, 5), (8, 12)])
def test_interleaved_pieces_assemble_to_reference(world, chunk):
    """The ranks emulated one after the other on one GPU, interleaved ownership
    (chunk c of 2^chunk rows belongs to rank c % world)."""
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = H.graded_square()
    ref = orc.stiffness_coo(c, e).tocsr()
    n = c.shape[0]
    with DeviceMesh(c, e) as mesh:
        _, _, whole = mesh.stiffness_csr()
    whole = whole.cpu().numpy()
    seen = 0
    for rank in range(world):
        with DeviceMesh(c, e, interleave=(world, rank, chunk)) as mesh:
            ip, ix, data = mesh.stiffness_csr()
            rows = mesh.global_rows().cpu().numpy()
        ip, ix, data = ip.cpu().numpy(), ix.cpu().numpy(), data.cpu().numpy()
        assert ip.shape[0] == rows.shape[0] + 1
        assert np.array_equal(rows, pdist.interleaved_rows(n, world, rank, chunk).numpy())
        piece = ref[rows]
        assert np.array_equal(ip, piece.indptr) and np.array_equal(ix, piece.indices)
        # bit-identical to the single-GPU values, whatever the partition
        lens = np.diff(ref.indptr)[rows]
        src = np.repeat(ref.indptr[rows] - piece.indptr[:-1], lens) + np.arange(ix.shape[0])
        assert np.array_equal(data, whole[src])
        seen += rows.shape[0]
    assert seen == n


@pytest.mark.parametrize('world', [2, 3, 8])
def test_row_ranges_concatenate_to_reference(world):
    """The ranks emulated one after the other on one GPU: each assembles its
    row range; concatenated pieces == reference CSR, independent of `world`."""
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = H.graded_square()
    ref = orc.stiffness_coo(c, e).tocsr()
    n = c.shape[0]
    ptr, idx, val = [np.zeros(1, dtype=np.int64)], [], []
    for lo, hi in pdist.partition_rows(n, world):
        with DeviceMesh(c, e, rows=(lo, hi)) as mesh:
            ip, ix, data = mesh.stiffness_csr()
            ptr.append(ip.cpu().numpy()[1:].astype(np.int64) + ptr[-1][-1])
            idx.append(ix.cpu().numpy())
            val.append(data.cpu().numpy())
    indptr = np.concatenate(ptr)
    assert np.array_equal(indptr, ref.indptr)
    assert np.array_equal(np.concatenate(idx), ref.indices)
    got = np.concatenate(val)
    assert np.abs(got - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
    # bit-identical to the single-GPU result: the summation order per row does
    # not depend on the partition
    with DeviceMesh(c, e) as mesh:
        _, _, whole = mesh.stiffness_csr()
    assert np.array_equal(got, whole.cpu().numpy())


def test_balanced_partition():
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = meshgen.unit_square(8)
    with DeviceMesh(c, e) as mesh:
        ip, _, _ = mesh.stiffness_csr()
        indptr = ip.cpu().numpy().astype(np.int64)
        for world in (1, 2, 4, 8):
            b = mesh.balanced_rows(world)
            assert b[0] == 0 and b[-1] == c.shape[0] and all(x <= y for x, y in zip(b, b[1:]))
            share = np.diff(indptr[b])
            assert share.max() <= 1.05 * share.mean() + 64, (world, share)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, result_path):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        c, e, b = meshgen.unit_square(6)
        c = meshgen.perturb_interior(c, b, 0.002)
        dm = pdist.DistributedMesh(c, e, device=rank, chunk_log2=6)
        ip, ix, data = dm.stiffness_csr()
        # where my chunks start in the global CSR (two kernels around one all-gather)
        full = orc.stiffness_coo(c, e).tocsr()
        offs, total = dm.global_chunk_offsets(ip)
        mine = np.arange(rank, (c.shape[0] + 63) // 64, world)
        assert int(total) == full.nnz
        assert np.array_equal(offs.cpu().numpy()[:mine.size], full.indptr[mine * 64])
        gptr, _ = dm.global_row_pointer(ip)
        assert np.array_equal(gptr.cpu().numpy(), full.indptr[dm.global_rows().cpu().numpy()])
        # the estimator by element blocks
        allb = np.vstack(b)
        from p1afempy_b200.device import to_device_indices
        d = to_device_indices(allb, rank, dm.local.ctx)
        none = torch.zeros((0, 2), dtype=torch.int32, device='cuda')
        x = H.u_nodal(c)
        fc = torch.from_numpy(H.f_load((c[e[:, 0]] + c[e[:, 1]] + c[e[:, 2]]) / 3.)).cuda()
        eta = dm.eta_r(torch.from_numpy(x).cuda(), d, none, None, fc).cpu().numpy()
        m0, m1 = dm.element_block()
        ref_eta = orc.eta_r(x, c, e, allb, np.zeros((0, 2), dtype=int),
                            lambda r: H.f_load((c[e[:, 0]] + c[e[:, 1]] + c[e[:, 2]]) / 3.),
                            H.g_neu)
        assert np.array_equal(eta, ref_eta[m0:m1])
        # load vectors by element blocks + one all-reduce: complete on every rank
        from p1afempy_b200 import cubature
        w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.CONICAL_GAUSS_4X4)
        xq = dm.local.quadrature_points(p).cpu().numpy()
        fq = np.stack([H.f_load(xq[q]) for q in range(w.shape[0])])
        ref_b = orc.load_vector_cubature(c, e, H.f_load, w, p)
        scale = 1e-12 * np.abs(ref_b).max()
        got_b = dm.load_vector_reduced(torch.from_numpy(fq).cuda(), w, p).cpu().numpy()
        assert got_b.shape == ref_b.shape and np.abs(got_b - ref_b).max() <= scale
        got_b = dm.load_vector_reduced(torch.from_numpy(fq[:, m0:m1].copy()).cuda(), w, p)
        assert np.abs(got_b.cpu().numpy() - ref_b).max() <= scale
        ref_b0 = orc.load_vector_midpoint(c, e, H.f_load)
        got_b0 = dm.load_vector_midpoint_reduced(fc).cpu().numpy()
        # (fc above is f at (c0 + c1 + c2) / 3, the oracle evaluates f at the reference's
        # c0 + (d21 + d31) / 3: equal to rounding only, also with one rank)
        assert np.abs(got_b0 - ref_b0).max() <= 1e-12 * np.abs(ref_b0).max()
        out = dm.gather(ip, ix, data, root=0)
        if rank == 0:
            ref = orc.stiffness_coo(c, e).tocsr()
            gi, gx, gd = (t.cpu().numpy() for t in out)
            ok = (np.array_equal(gi, ref.indptr) and np.array_equal(gx, ref.indices)
                  and np.abs(gd - ref.data).max() <= 1e-12 * np.abs(ref.data).max())
            with open(result_path, 'w') as fh:
                fh.write('ok' if ok else 'mismatch')
        dm.close()
    finally:
        dist.destroy_process_group()


def test_one_rank_nccl(tmp_path):
    """The NCCL code path with a world of one (any box): DistributedMesh, the global
    row pointer kernels, the estimator by element blocks, the gather."""
    import torch.multiprocessing as mp
    result = tmp_path / 'r.txt'
    mp.spawn(_nccl_worker, args=(1, _free_port(), str(result)), nprocs=1, join=True)
    assert result.read_text() == 'ok'


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_ranks_nccl(tmp_path):
    import torch.multiprocessing as mp
    result = tmp_path / 'r.txt'
    mp.spawn(_nccl_worker, args=(2, _free_port(), str(result)), nprocs=2, join=True)
    assert result.read_text() == 'ok'


def test_device_mesh_generator_matches_host():
    """meshgen.unit_square_device == meshgen.unit_square (== the reference's
    refineNVB, tests/test_meshgen_vs_reference.py) bit for bit."""
    c, e, b = meshgen.unit_square(8)
    ct, et, bt = meshgen.unit_square_device(8, host_levels=3)
    assert np.array_equal(ct.cpu().numpy(), c)
    assert np.array_equal(et.cpu().numpy(), e)
    assert np.array_equal(bt[0].cpu().numpy(), b[0])


def test_device_refine_nvb_matches_host():
    """meshgen.refine_nvb_device == meshgen.refine_nvb (== the reference's
    refineNVB) on randomly marked L-shape and square meshes."""
    rng = np.random.default_rng(5)
    for start in (meshgen.l_shape, meshgen.criss_square):
        c, e, b = start()
        ct = torch.from_numpy(c).cuda()
        et = torch.from_numpy(e.astype(np.int32)).cuda()
        bt = [torch.from_numpy(x.astype(np.int32)).cuda() for x in b]
        for step in range(9):
            marked = np.flatnonzero(rng.random(e.shape[0]) < (1.0 if step < 2 else 0.3))
            c, e, b = meshgen.refine_nvb(c, e, marked, b)
            ct, et, bt = meshgen.refine_nvb_device(
                ct, et, torch.from_numpy(marked).cuda(), bt)
            assert np.array_equal(ct.cpu().numpy(), c)
            assert np.array_equal(et.cpu().numpy(), e)
            for x, y in zip(bt, b):
                assert np.array_equal(x.cpu().numpy(), y)


@pytest.mark.parametrize('world,chunk', [(2, 5), (4, 3)])
def test_every_operator_on_interleaved_shares(world, chunk):
    """Mass, generalised stiffness, weighted mass (plan route: P1_OP_LOCAL; cold
    route: inside the scatter) and both load vectors on the rows of every rank
    (emulated one after the other) == the one-GPU result restricted to those rows,
    bit for bit."""
    from p1afempy_b200 import _lib, cubature
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = H.graded_square()
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.DAYTAYLOR)
    ct, et = torch.from_numpy(c).cuda(), torch.from_numpy(e.astype(np.int32)).cuda()
    u = torch.from_numpy(H.u_nodal(c)).cuda()
    with DeviceMesh(ct, et) as mesh:
        xq = mesh.quadrature_points(p)
        qs = [torch.from_numpy(np.stack([f(xq[q].cpu().numpy()) for q in range(len(w))])).cuda()
              for f in (H.a11, H.a12, H.a21, H.a22)]
        fq = torch.from_numpy(np.stack([H.f_load(xq[q].cpu().numpy())
                                        for q in range(len(w))])).cuda()
        fc = torch.from_numpy(H.f_load(mesh.centroids().cpu().numpy())).cuda()
        phiq = (1. + mesh.nodal_at_quadrature(u, p) ** 2).contiguous()
        whole = {'mass': mesh.mass_csr(), 'general': mesh.general_stiffness_csr(*qs, w),
                 'weighted': mesh.weighted_mass_csr(phiq, w, p)}
        whole = {k: [t.cpu().numpy() for t in v] for k, v in whole.items()}
        b_full = mesh.load_vector(fq, w, p).cpu().numpy()
        b0_full = mesh.load_vector_midpoint(fc).cpu().numpy()
    ref_g = orc.general_stiffness_coo(c, e, H.a11, H.a12, H.a21, H.a22, w, p).tocsr()
    H.values_close_by_row(whole['general'][2], ref_g, 1e-12)
    for rank in range(world):
        with DeviceMesh(ct, et, interleave=(world, rank, chunk)) as mesh:
            rows = mesh.global_rows().cpu().numpy()
            got = {'mass': mesh.mass_csr(), 'general': mesh.general_stiffness_csr(*qs, w),
                   'weighted': mesh.weighted_mass_csr(phiq, w, p)}
            w_args = _lib.QuadArgs(None, None, None, None, phiq.data_ptr(), w.ctypes.data,
                                   p.ctypes.data, len(w))
            g_args = _lib.QuadArgs(qs[0].data_ptr(), qs[1].data_ptr(), qs[2].data_ptr(),
                                   qs[3].data_ptr(), None, w.ctypes.data, None, len(w))
            cold = {'mass': mesh.assemble_cold(_lib.OP_MASS).wait(),
                    'general': mesh.assemble_cold(_lib.OP_GENERAL, quad=g_args).wait(),
                    'weighted': mesh.assemble_cold(_lib.OP_WEIGHTED, quad=w_args).wait()}
            for name, (ip, ix, data) in list(got.items()) + list(cold.items()):
                fip, fix, fdata = whole[name]
                lens = np.diff(fip)[rows]
                src = (np.repeat(fip[rows] - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
                       + np.arange(int(lens.sum())))
                assert np.array_equal(np.diff(ip.cpu().numpy()), lens), name
                assert np.array_equal(ix.cpu().numpy(), fix[src]), name
                assert np.array_equal(data.cpu().numpy(), fdata[src]), name
            assert np.array_equal(mesh.load_vector(fq, w, p).cpu().numpy(), b_full[rows])
            assert np.array_equal(mesh.load_vector_midpoint(fc).cpu().numpy(), b0_full[rows])


@pytest.mark.parametrize('world', [3, 8])
def test_load_vector_by_element_blocks(world):
    """What DistributedMesh.load_vector_reduced adds up: the vectors of contiguous
    element blocks over all N nodes.  Their sum in rank order agrees with the
    reference to rounding (sums associated by block), each block's vector is the
    reference's for that block bit for bit."""
    from p1afempy_b200 import cubature
    from p1afempy_b200.device import DeviceMesh
    c, e, _ = H.perturbed_square(5)
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.CONICAL_GAUSS_4X4)
    ct, et = torch.from_numpy(c).cuda(), torch.from_numpy(e.astype(np.int32)).cuda()
    with DeviceMesh(ct, et) as mesh:
        xq = mesh.quadrature_points(p).cpu().numpy()
    fq = np.stack([H.f_load(xq[q]) for q in range(w.shape[0])])
    ref = orc.load_vector_cubature(c, e, H.f_load, w, p)
    m, n = e.shape[0], c.shape[0]
    total = np.zeros(n)
    for r in range(world):
        m0, m1 = m * r // world, m * (r + 1) // world
        with DeviceMesh(ct, et[m0:m1], n_nodes=n, check=False) as part:
            b = part.load_vector(torch.from_numpy(fq[:, m0:m1].copy()).cuda(), w, p).cpu().numpy()
        assert b.shape == (n,)
        assert np.array_equal(b, orc.load_vector_cubature(c, e[m0:m1], H.f_load, w, p))
        total += b
    assert np.abs(total - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize('world', [2, 5])
def test_eta_r_by_element_blocks(world):
    """Multi-GPU estimator: every rank (emulated) computes eta_R of its element block
    on the sub-mesh of the elements touching the block's nodes; the pieces in rank
    order are the one-GPU vector bit for bit -- with Dirichlet and Neumann parts."""
    from p1afempy_b200.device import DeviceMesh, to_device_indices
    for c, e, b in (H.graded_square(), H.perturbed_square(5)):
        allb = np.vstack(b)
        cut = (3 * allb.shape[0]) // 5
        x = H.u_nodal(c) + 0.1 * c[:, 0]
        ref = orc.eta_r(x, c, e, allb[:cut], allb[cut:], H.f_load, H.g_neu)
        with DeviceMesh(c, e) as mesh:
            d = to_device_indices(allb[:cut], mesh.device, mesh.ctx)
            n = to_device_indices(allb[cut:], mesh.device, mesh.ctx)
            length, mid = mesh.boundary_edges(n)
            term = length * torch.from_numpy(H.g_neu(mid.cpu().numpy())).cuda()
            fc = torch.from_numpy(H.f_load(mesh.centroids().cpu().numpy())).cuda()
            xt = torch.from_numpy(x).cuda()
            whole = mesh.eta_r(xt, d, n, term, fc).cpu().numpy()
            assert np.array_equal(whole, ref)
            m = e.shape[0]
            pieces = [mesh.eta_r_block(xt, d, n, term, fc, m * r // world, m * (r + 1) // world)
                      for r in range(world)]
            assert np.array_equal(torch.cat(pieces).cpu().numpy(), ref)
