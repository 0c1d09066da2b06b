"""Host-side arithmetic that the device code mirrors (no GPU needed)."""
import numpy as np
import pytest

from p1afempy_b200 import meshgen
from p1afempy_b200.device import interleaved_count, scipy_index_dtype


@pytest.mark.parametrize('n', [0, 1, 63, 64, 65, 1000, 4097, 12345])
@pytest.mark.parametrize('world', [1, 2, 3, 8])
@pytest.mark.parametrize('chunk_log2', [0, 4, 6, 12])
def test_interleaved_count_matches_dealing_chunks(n, world, chunk_log2):
    """plan.cuh Own::count / device.interleaved_count: chunk c of 2^k rows belongs to
    rank c % world."""
    owner = (np.arange(n) >> chunk_log2) % world
    counts = [interleaved_count(n, world, r, chunk_log2) for r in range(world)]
    assert counts == [int((owner == r).sum()) for r in range(world)]
    assert sum(counts) == n


def test_scipy_index_dtype_rule():
    """scipy/sparse/_coo.py:419: tocsr picks the index dtype from max(nnz_coo, N)."""
    assert scipy_index_dtype(1, 5, 5) == np.int32
    assert scipy_index_dtype((2 ** 31 - 1) // 9, 5, 5) == np.int32
    assert scipy_index_dtype((2 ** 31 - 1) // 9 + 1, 5, 5) == np.int64
    assert scipy_index_dtype(1, 2 ** 31, 5) == np.int64


@pytest.mark.parametrize('level', [0, 2, 4])
def test_cold_capacity_bounds_nnz_on_manifold_meshes(level):
    """The cold path's output capacity 2 N + 3 M >= nnz = N + 2 E whenever every boundary
    node has one open fan (E = (3 M + B) / 2 with B <= N boundary edges)."""
    c, e, b = meshgen.unit_square(level)
    m, n = e.shape[0], c.shape[0]
    edges = np.sort(np.vstack([e[:, [0, 1]], e[:, [1, 2]], e[:, [2, 0]]]), axis=1)
    n_edges = np.unique(edges, axis=0).shape[0]
    assert n + 2 * n_edges <= 2 * n + 3 * m
    assert sum(x.shape[0] for x in b) == 2 * n_edges - 3 * m     # boundary edges
