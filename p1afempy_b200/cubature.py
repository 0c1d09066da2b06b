"""Cubature rules on the reference triangle (0,0),(1,0),(0,1) -- rule as DATA.

The reference looks its rules up in the third-party ``triangle-cubature``
package (``get_rule(rule).weights_and_integration_points`` with ``.weights``
(n_qp,) and ``.integration_points`` (n_qp, 2) as (eta, xi); weights sum to
1/2; call sites /root/reference/p1afempy/solvers.py:88-89, 316-317, 568-569).
That package is not vendored by the reference, not pinned, and not available
offline, so this module

* accepts ANY of: a member of the real ``triangle_cubature`` enum (resolved
  through that package when it is importable), a member of the enum below,
  an object exposing ``weights``/``integration_points`` (optionally behind a
  ``weights_and_integration_points`` attribute), or a ``(weights, points)``
  pair;
* ships MIDPOINT (one point (1/3,1/3), weight 1/2 -- the only table the
  reference's tests pin numerically, test_solver.py:88-129), the vertex rule
  LAUFFER_LINEAR, and two rules under their own names: EDGE_MIDPOINTS (degree
  2) and CONICAL_GAUSS_4X4 (degree 6: the 4x4 Gauss-Legendre rule on the unit
  square collapsed onto the triangle by (s, t) -> (s, t(1-s)) with weight
  factor (1-s); 16 points, positive weights, sum 1/2);
* resolves the names DAYTAYLOR and SMPLX1 through the real ``triangle_cubature``
  package when it is importable; when it is not, they fall back to
  CONICAL_GAUSS_4X4 / EDGE_MIDPOINTS and say so once with a
  ``StandInRuleWarning``: same degree of exactness, but NOT the numeric table
  the reference would use ("parity unpinned"), so
  results for non-polynomial integrands differ from the reference's in the
  quadrature error.  Every parity test feeds the oracle and the CUDA path the
  same arrays, so parity is unaffected.
"""
from __future__ import annotations

import enum
import warnings
from dataclasses import dataclass

import numpy as np


class CubatureRuleEnum(enum.Enum):
    MIDPOINT = 1
    LAUFFER_LINEAR = 2
    SMPLX1 = 3
    DAYTAYLOR = 4
    EDGE_MIDPOINTS = 5
    CONICAL_GAUSS_4X4 = 6


class StandInRuleWarning(UserWarning):
    """A rule name of ``triangle_cubature`` was resolved to this module's stand-in
    table because that package is not installed."""


_STAND_INS = {CubatureRuleEnum.DAYTAYLOR: CubatureRuleEnum.CONICAL_GAUSS_4X4,
              CubatureRuleEnum.SMPLX1: CubatureRuleEnum.EDGE_MIDPOINTS}
_warned = set()


@dataclass(frozen=True)
class WeightsAndIntegrationPoints:
    weights: np.ndarray
    integration_points: np.ndarray


@dataclass(frozen=True)
class CubatureRule:
    weights_and_integration_points: WeightsAndIntegrationPoints
    degree_of_exactness: int
    name: str


def _conical_gauss(n: int) -> WeightsAndIntegrationPoints:
    x, w = np.polynomial.legendre.leggauss(n)
    s = 0.5 * (x + 1.0)
    ws = 0.5 * w
    pts = []
    wts = []
    for i in range(n):
        for j in range(n):
            pts.append((s[i], s[j] * (1.0 - s[i])))
            wts.append(ws[i] * ws[j] * (1.0 - s[i]))
    return WeightsAndIntegrationPoints(
        weights=np.asarray(wts, dtype=np.float64),
        integration_points=np.asarray(pts, dtype=np.float64))


_TABLES = {
    CubatureRuleEnum.MIDPOINT: CubatureRule(
        WeightsAndIntegrationPoints(
            weights=np.array([0.5]),
            integration_points=np.array([[1. / 3., 1. / 3.]])),
        degree_of_exactness=1, name='midpoint'),
    CubatureRuleEnum.LAUFFER_LINEAR: CubatureRule(
        WeightsAndIntegrationPoints(
            weights=np.array([1. / 6., 1. / 6., 1. / 6.]),
            integration_points=np.array([[0., 0.], [1., 0.], [0., 1.]])),
        degree_of_exactness=1, name='lauffer-linear (vertices)'),
    CubatureRuleEnum.EDGE_MIDPOINTS: CubatureRule(
        WeightsAndIntegrationPoints(
            weights=np.array([1. / 6., 1. / 6., 1. / 6.]),
            integration_points=np.array([[.5, 0.], [.5, .5], [0., .5]])),
        degree_of_exactness=2, name='edge midpoints'),
    CubatureRuleEnum.CONICAL_GAUSS_4X4: CubatureRule(
        _conical_gauss(4), degree_of_exactness=6, name='4x4 conical Gauss'),
}


def _real_rule(name):
    """The table ``triangle_cubature`` ships under `name`, or None offline."""
    try:
        from triangle_cubature.cubature_rule import CubatureRuleEnum as Real
        from triangle_cubature.rule_factory import get_rule as real_get_rule
        if real_get_rule is get_rule:   # the oracle's import shim re-exports this module
            return None
        return real_get_rule(rule=Real[name])
    except Exception:
        return None


def get_rule(rule: CubatureRuleEnum) -> CubatureRule:
    """Same call shape as ``triangle_cubature.rule_factory.get_rule``."""
    rule = CubatureRuleEnum(rule)
    if rule in _STAND_INS:
        real = _real_rule(rule.name)
        if real is not None:
            return real
        if rule not in _warned:
            _warned.add(rule)
            warnings.warn(
                'triangle_cubature is not installed: %s resolves to the stand-in rule %s (same '
                'degree of exactness, different points and weights than the reference uses)'
                % (rule.name, _STAND_INS[rule].name), StandInRuleWarning, stacklevel=3)
        rule = _STAND_INS[rule]
    return _TABLES[rule]


def resolve_rule(rule) -> tuple[np.ndarray, np.ndarray]:
    """Turn whatever the caller passed as ``cubature_rule`` into
    ``(weights (n_qp,), points (n_qp, 2))`` float64 C-contiguous arrays."""
    if isinstance(rule, CubatureRuleEnum):
        wip = get_rule(rule).weights_and_integration_points
    elif isinstance(rule, enum.Enum):
        # a member of the real triangle_cubature enum
        from triangle_cubature.rule_factory import get_rule as real_get_rule
        wip = real_get_rule(rule=rule).weights_and_integration_points
    elif hasattr(rule, 'weights_and_integration_points'):
        wip = rule.weights_and_integration_points
    elif hasattr(rule, 'weights') and hasattr(rule, 'integration_points'):
        wip = rule
    else:
        w, p = rule
        wip = WeightsAndIntegrationPoints(np.asarray(w), np.asarray(p))
    w = np.ascontiguousarray(wip.weights, dtype=np.float64).reshape(-1)
    p = np.ascontiguousarray(wip.integration_points,
                             dtype=np.float64).reshape(-1, 2)
    if w.shape[0] != p.shape[0]:
        raise ValueError('cubature rule: weights/points length mismatch')
    return w, p
