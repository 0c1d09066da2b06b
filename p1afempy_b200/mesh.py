"""Drop-in for the three helpers of ``p1afempy.mesh`` that sit on the hot path
(/root/reference/p1afempy/mesh.py:10-51, 87-161)."""
from __future__ import annotations

from .device import DeviceMesh, to_host

__all__ = ['get_area', 'get_directional_vectors', 'provide_geometric_data',
           'get_element_to_neighbours']


def get_area(coordinates, elements):
    """Signed element areas (mesh.py:10-28)."""
    with DeviceMesh(coordinates, elements) as mesh:
        area, _, _ = mesh.geometry(area=True, vectors=False)
        return to_host(area)


def get_directional_vectors(coordinates, elements):
    """(d21, d31): vertex0->vertex1 and vertex0->vertex2 (mesh.py:31-51)."""
    with DeviceMesh(coordinates, elements) as mesh:
        _, d21, d31 = mesh.geometry(area=False, vectors=True)
        return to_host(d21), to_host(d31)


def provide_geometric_data(elements, boundaries):
    """(element2edges, edge2nodes, boundaries_to_edges) with the reference's
    edge numbering (mesh.py:87-161); int64 like the reference.  Raises
    ValueError when the boundary edges are not all listed (the reference fails
    with a shape mismatch at mesh.py:150)."""
    with DeviceMesh(None, elements) as mesh:
        e2e, e2n, b2e = mesh.geometric_data(list(boundaries))
        return to_host(e2e), to_host(e2n), [to_host(b) for b in b2e]


def get_element_to_neighbours(elements):
    """element2neighbours[k, i] = the element sharing edge i (vertex i -> vertex i+1)
    of element k, -1 on the boundary (mesh.py:495-528; SURVEY.md section 8(f), rank
    4).  From the half-edge twins of the assembly path's node fans; int64."""
    with DeviceMesh(None, elements) as mesh:
        return to_host(mesh.element_neighbours())
