"""Mesh I/O: the reference's text readers (/root/reference/p1afempy/io_helpers.py:5-132,
same names and arguments) plus binary ``.npy`` / ``.npz`` storage -- text files are
unusable from a million triangles on.  Host code:
files are read on the host whatever computes afterwards."""
from __future__ import annotations

from pathlib import Path

import numpy as np

__all__ = ['read_boundary_condition', 'read_coordinates', 'read_elements', 'read_mesh',
           'save_mesh', 'load_mesh']


def read_boundary_condition(path_to_boundary: Path, shift_indices: bool = False) -> np.ndarray:
    """(K, 2) uint32 boundary edges (io_helpers.py:5-39); a single edge keeps two
    dimensions; shift_indices: i := i - 1 for 1-based files."""
    boundary_indices = np.loadtxt(path_to_boundary).astype(np.uint32)
    if shift_indices:
        boundary_indices = boundary_indices - 1
    if len(boundary_indices.shape) == 1:
        return boundary_indices[None, :]
    return boundary_indices


def read_coordinates(path_to_coordinates: Path) -> np.ndarray:
    """(N, 2) float64 (io_helpers.py:42-63)."""
    return np.loadtxt(path_to_coordinates)


def read_elements(path_to_elements: Path, shift_indices: bool = False) -> np.ndarray:
    """(M, 3) uint32 (io_helpers.py:66-94)."""
    elements = np.loadtxt(path_to_elements).astype(np.uint32)
    if shift_indices:
        elements = elements - 1
    return elements


def read_mesh(path_to_coordinates: Path, path_to_elements: Path,
              shift_indices: bool = False) -> tuple[np.ndarray, np.ndarray]:
    """(coordinates, elements) (io_helpers.py:97-132)."""
    return (read_coordinates(path_to_coordinates=path_to_coordinates),
            read_elements(path_to_elements=path_to_elements, shift_indices=shift_indices))


def save_mesh(path, coordinates, elements, boundaries=()):
    """One ``.npz`` file: coordinates float64 (N, 2), elements int32 (M, 3) (int64 when
    an index needs it), boundary_0 ... boundary_k (K_j, 2).  Accepts numpy arrays and
    torch tensors (CUDA tensors are copied to the host)."""
    def host(a):
        return a.detach().cpu().numpy() if hasattr(a, 'detach') else np.asarray(a)
    e = host(elements)
    kind = np.int32 if e.size == 0 or int(e.max()) <= np.iinfo(np.int32).max else np.int64
    arrays = {'coordinates': np.ascontiguousarray(host(coordinates), dtype=np.float64),
              'elements': np.ascontiguousarray(e, dtype=kind)}
    for j, b in enumerate(boundaries):
        arrays['boundary_%d' % j] = np.ascontiguousarray(host(b).reshape(-1, 2), dtype=kind)
    path = Path(path)
    with open(path if path.suffix == '.npz' else path.with_suffix('.npz'), 'wb') as fh:
        np.savez(fh, **arrays)


def load_mesh(path):
    """-> (coordinates, elements, [boundaries]) as written by save_mesh."""
    path = Path(path)
    with np.load(path if path.suffix == '.npz' else path.with_suffix('.npz')) as z:
        n_b = sum(1 for k in z.files if k.startswith('boundary_'))
        return (z['coordinates'], z['elements'], [z['boundary_%d' % j] for j in range(n_b)])
