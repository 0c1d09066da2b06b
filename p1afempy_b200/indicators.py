"""Drop-in for ``p1afempy.indicators.compute_eta_r``
(/root/reference/p1afempy/indicators.py:7-86)."""
from __future__ import annotations

import numpy as np

from .device import DeviceMesh, to_device_indices, to_host
from .solvers import _values_on_host

__all__ = ['compute_eta_r']


def compute_eta_r(x, coordinates, elements, dirichlet, neumann, f, g):
    """Squared residual indicators per element.  The callbacks ``f`` (at the
    element centroids) and ``g`` (at the Neumann edge midpoints) run on the
    host, everything else on the device."""
    with DeviceMesh(coordinates, elements) as mesh:
        d = to_device_indices(np.asarray(dirichlet).reshape(-1, 2), mesh.device,
                              mesh.ctx)
        n = to_device_indices(np.asarray(neumann).reshape(-1, 2), mesh.device,
                              mesh.ctx)
        term = None
        if n.shape[0] > 0:
            length, mid = mesh.boundary_edges(n)
            gm = _values_on_host(g, to_host(mid), n.shape[0])
            term = to_host(length) * gm
        fc = _values_on_host(f, to_host(mesh.centroids()), mesh.n_elements)
        return to_host(mesh.eta_r(x, d, n, term, fc))
