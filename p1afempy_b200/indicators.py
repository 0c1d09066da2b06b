"""Drop-in for ``p1afempy.indicators.compute_eta_r``
(/root/reference/p1afempy/indicators.py:7-86)."""
from __future__ import annotations

import numpy as np
import torch

from .device import DeviceMesh, to_device_indices, to_host
from .solvers import _callback_at

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
            gm = _callback_at(g, mid)
            # |E| * g(midpoint), indicators.py:74-76 (one multiplication either way)
            term = (length * gm) if isinstance(gm, torch.Tensor) else to_host(length) * gm
        fc = _callback_at(f, mesh.centroids())
        return to_host(mesh.eta_r(x, d, n, term, fc))
