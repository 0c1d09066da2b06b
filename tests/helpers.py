"""Shared test inputs: meshes, callbacks, tolerances."""
import os

import numpy as np

from p1afempy_b200 import meshgen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def perturbed_square(k, seed=0):
    c, e, b = meshgen.unit_square(k)
    c = meshgen.perturb_interior(c, b, 0.2 * 2.0 ** (-k), seed=seed)
    return c, e, b


def graded_square(k_uniform=2, n_adaptive=4, seed=0):
    """Locally refined, perturbed unit square: mixed valences, hanging-node
    free (NVB closure), several refinement levels side by side."""
    c, e, b = meshgen.unit_square(k_uniform)
    rng = np.random.default_rng(seed)
    for _ in range(n_adaptive):
        cent = (c[e[:, 0]] + c[e[:, 1]] + c[e[:, 2]]) / 3.
        near = np.flatnonzero(np.hypot(cent[:, 0] - 0.3, cent[:, 1] - 0.6)
                              < 0.25)
        extra = rng.choice(e.shape[0], size=max(1, e.shape[0] // 20),
                           replace=False)
        c, e, b = meshgen.refine_nvb(c, e, np.union1d(near, extra), b)
    h = np.sqrt(np.min(np.abs(
        0.5 * ((c[e[:, 1], 0] - c[e[:, 0], 0]) * (c[e[:, 2], 1] - c[e[:, 0], 1])
               - (c[e[:, 1], 1] - c[e[:, 0], 1]) * (c[e[:, 2], 0] - c[e[:, 0], 0])))))
    c = meshgen.perturb_interior(c, b, 0.1 * h, seed=seed)
    return c, e, b


# coefficient callbacks of benchmark config C3
def a11(r): return -(1. + r[:, 0] ** 2)
def a12(r): return -0.25 * r[:, 0] * r[:, 1]
def a21(r): return -0.5 * r[:, 0] * r[:, 1]
def a22(r): return -(2. + np.sin(np.pi * r[:, 1]))
def phi(u): return 1. + u ** 2
def f_load(r): return np.sin(2. * np.pi * r[:, 0]) * np.cos(3. * np.pi * r[:, 1]) + 1.
def g_neu(r): return np.cos(r[:, 0]) - 0.5 * r[:, 1]
def f_nl(x): return 1.7 * x ** 2 + 0.3 * x ** 3
def u_nodal(c): return np.sin(np.pi * c[:, 0]) * np.sin(np.pi * c[:, 1])


def values_close_by_row(data, ref, rtol=1e-12):
    """The value part of the parity check: |v - v_ref| <= rtol * max|row of ref|."""
    counts = np.diff(ref.indptr)
    rowmax = np.zeros(ref.shape[0])
    nz = counts > 0
    rowmax[nz] = np.maximum.reduceat(np.abs(ref.data), ref.indptr[:-1][nz])
    scale = np.repeat(rowmax, counts)
    err = np.abs(data - ref.data)
    bad = err > rtol * scale
    assert not bad.any(), (int(bad.sum()), float((err / np.maximum(scale, 1e-300)).max()))


def csr_values_close(a, ref, rtol=1e-12):
    """SURVEY.md section 8(d) parity check: identical index arrays (dtype and
    shape included), values within rtol * max|row of ref|."""
    assert a.shape == ref.shape, (a.shape, ref.shape)
    assert a.indptr.dtype == ref.indptr.dtype, (a.indptr.dtype, ref.indptr.dtype)
    assert a.indices.dtype == ref.indices.dtype
    assert np.array_equal(a.indptr, ref.indptr)
    assert np.array_equal(a.indices, ref.indices)
    assert a.data.dtype == np.float64
    counts = np.diff(ref.indptr)
    rowmax = np.zeros(ref.shape[0])
    nz = counts > 0
    rowmax[nz] = np.maximum.reduceat(np.abs(ref.data), ref.indptr[:-1][nz])
    scale = np.repeat(rowmax, counts)
    err = np.abs(a.data - ref.data)
    bad = err > rtol * scale
    assert not bad.any(), (int(bad.sum()), float((err / np.maximum(scale, 1e-300)).max()))
