"""Doerfler marking for the solve -- estimate -- MARK -- refine loop (SURVEY.md section
8(f), rank 4).  The reference has no marking routine (``adaptiveAlgorithm.m`` is
unticked in its README); this is the criterion of the p1afem package it adapts: the
smallest set of elements, taken in order of decreasing indicator (stable: ties by
index), whose indicators sum to at least ``theta`` times the total."""
from __future__ import annotations

import numpy as np
import torch

__all__ = ['doerfler_marking', 'doerfler_marking_device']


def doerfler_marking_device(eta: torch.Tensor, theta: float = 0.5) -> torch.Tensor:
    """eta: (M,) float64 CUDA tensor of squared indicators -> int64 CUDA tensor with
    the indices of the marked elements, in order of decreasing indicator.  Sort,
    prefix sum and search run on the device (torch: the one library sort of the
    package); nothing but the count comes back to the host."""
    if eta.numel() == 0:
        return torch.zeros(0, dtype=torch.int64, device=eta.device)
    order = torch.sort(eta, descending=True, stable=True).indices
    csum = torch.cumsum(eta[order], 0)
    k = int(torch.searchsorted(csum, float(theta) * csum[-1])) + 1
    return order[:min(k, eta.numel())].contiguous()


def doerfler_marking(eta, theta: float = 0.5) -> np.ndarray:
    """numpy in, numpy out (int64 element indices); the work happens on the device."""
    from .device import _require_cuda
    _require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(eta, dtype=np.float64))).cuda()
    return doerfler_marking_device(t, theta).cpu().numpy()
