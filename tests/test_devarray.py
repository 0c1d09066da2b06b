"""The ndarray stand-in that lets user callbacks run on device data (CPU
tensors here: the protocol is device-agnostic)."""
import numpy as np
import torch

from p1afempy_b200 import devarray
import helpers as H
import helpers_laplace as L


def test_reference_style_callbacks_run_on_the_proxy():
    pts = np.random.default_rng(0).random((257, 2))
    t = torch.from_numpy(pts)
    fns = [H.a11, H.a12, H.a21, H.a22, H.f_load, H.g_neu, L.f, L.u, L.uD,
           lambda r: -np.ones(r.shape[0]), lambda r: np.zeros(r.shape[0]), lambda r: 3.0,
           lambda r: np.where(r[:, 0] > 0.5, r[:, 1] ** 2, -r[:, 0]),
           lambda r: np.sqrt(np.abs(r[:, 0] - 0.3)) + np.exp(-r[:, 1]) / (1. + r[:, 0])]
    for fn in fns:
        got = devarray.evaluate(fn, t, pts.shape[0])
        assert got is not None and got.dtype == torch.float64 and got.shape == (257,)
        ref = np.broadcast_to(np.asarray(fn(pts), dtype=float), (257,))
        assert np.allclose(got.numpy(), ref, rtol=1e-14, atol=0)
    u = torch.from_numpy(pts[:, 0].copy())
    assert np.allclose(devarray.evaluate(H.phi, u, 257).numpy(), H.phi(pts[:, 0]))


def test_unsupported_callbacks_are_reported_not_guessed():
    t = torch.from_numpy(np.random.default_rng(1).random((16, 2)))
    assert devarray.evaluate(L.g, t, 16) is None                   # mask assignment
    assert devarray.evaluate(lambda r: np.array([x for x in r]), t, 16) is None
    assert devarray.evaluate(lambda r: np.asarray(r)[:, 0], t, 16) is None
