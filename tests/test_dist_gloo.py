"""Host-side logic of the multi-GPU path on CPU: world_size 2, gloo.  Each rank
owns a contiguous row range; the local pieces (here cut out of the oracle's
CSR, the CUDA path is exercised by tests/test_gpu_dist.py) must concatenate to
the reference matrix bit for bit, whatever the number of ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import p1_oracle as orc
from p1afempy_b200 import dist as pdist
from p1afempy_b200 import meshgen


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_path, layout):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        c, e, b = meshgen.unit_square(4)
        c = meshgen.perturb_interior(c, b, 0.01)
        ref = orc.stiffness_coo(c, e).tocsr()
        n = ref.shape[0]
        if layout == 'contiguous':
            lo, hi = pdist.partition_rows(n, world)[rank]
            rows = torch.arange(lo, hi, dtype=torch.int64)
        else:
            rows = pdist.interleaved_rows(n, world, rank, 5)
        piece = ref[rows.numpy()]
        indptr = torch.from_numpy(piece.indptr.astype(np.int32))
        indices = torch.from_numpy(piece.indices.astype(np.int32))
        data = torch.from_numpy(piece.data.copy())
        offset, total, every = pdist.nnz_offsets(data.numel())
        assert total == ref.nnz
        if layout == 'interleaved':
            # the global row pointer from the per-chunk nnz of every rank (one all-gather)
            gptr, gtotal = pdist.global_row_pointer(indptr, n, world, rank, 5)
            assert int(gtotal) == ref.nnz
            assert np.array_equal(gptr.numpy(), ref.indptr[rows.numpy()])
        out = pdist.gather_csr(indptr, indices, data, rows, n, root=0)
        if rank == 0:
            gi, gx, gd = (t.numpy() for t in out)
            ok = (np.array_equal(gi, ref.indptr) and np.array_equal(gx, ref.indices)
                  and np.array_equal(gd, ref.data))
            with open(result_path, 'w') as fh:
                fh.write('ok' if ok else 'mismatch')
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,layout', [(2, 'contiguous'), (2, 'interleaved'),
                                          (3, 'interleaved')])
def test_row_pieces_assemble_to_reference(tmp_path, world, layout):
    result = tmp_path / 'result.txt'
    mp.spawn(_worker, args=(world, _free_port(), str(result), layout), nprocs=world,
             join=True)
    assert result.read_text() == 'ok'


def test_interleaved_rows_cover_everything():
    for n in (1, 7, 100, 4097, 33562625):
        for w in (1, 2, 8):
            for c in (0, 3, 12):
                if n > 10 ** 6 and c < 12:
                    continue
                got = torch.cat([pdist.interleaved_rows(n, w, r, c) for r in range(w)])
                assert got.numel() == n and torch.equal(torch.sort(got).values,
                                                         torch.arange(n))


def test_partition_rows_covers_everything():
    for n in (0, 1, 7, 100, 33562625):
        for w in (1, 2, 4, 8):
            parts = pdist.partition_rows(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
