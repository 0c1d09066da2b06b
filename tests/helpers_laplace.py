"""Problem data of the reference's Laplace example, restated from
/root/reference/p1afempy/tests/auto/example_setup.py:4-36, and the pyramid
function of test_sanity_check.py:14-20 (vectorised)."""
import numpy as np

OMEGA = 7. / 4. * np.pi


def u(r):
    return np.sin(OMEGA * 2. * r[:, 0]) * np.sin(OMEGA * r[:, 1])


def f(r):
    return 5. * OMEGA ** 2 * np.sin(OMEGA * 2. * r[:, 0]) * np.sin(OMEGA * r[:, 1])


def uD(r):
    return u(r)


def g(r):
    """Neumann data; piecewise by boolean-mask assignment, i.e. a callback
    that can only be evaluated on the host."""
    out = np.zeros(r.shape[0])
    right = r[:, 0] == 1
    upper = r[:, 1] == 1
    out[right] = -2. * OMEGA * np.sin(OMEGA * r[right, 1]) * np.cos(OMEGA * 2. * r[right, 0])
    out[upper] = OMEGA * np.sin(OMEGA * 2. * r[upper, 0]) * np.cos(OMEGA * r[upper, 1])
    return out


def sanity_u(c):
    x, y = c[:, 0], c[:, 1]
    inside = (-1. < x) & (x < 1.) & (-1. < y) & (y < 1.)
    return np.where(inside, np.minimum(1. - np.abs(x), 1. - np.abs(y)), 0.)
