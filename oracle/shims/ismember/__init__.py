"""Stand-in for the third-party ``ismember`` package (absent offline).

Only used so that the real reference can be IMPORTED from /root/reference for
oracle validation (p1afempy/mesh.py:4, p1afempy/refinement.py:2).  No function
on the hot path calls into it.
"""
import numpy as np


def _row_keys(a, b):
    a = np.atleast_2d(np.asarray(a))
    b = np.atleast_2d(np.asarray(b))
    both = np.vstack([a, b])
    _, inv = np.unique(both, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    return inv[:a.shape[0]], inv[a.shape[0]:]


def ismember(a, b, method=None):
    """Returns (mask over a, index into b for every True of the mask)."""
    if method == 'rows':
        ka, kb = _row_keys(a, b)
    else:
        ka, kb = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    first = {}
    for pos, key in enumerate(kb.tolist()):
        first.setdefault(key, pos)
    mask = np.array([k in first for k in ka.tolist()], dtype=bool)
    idx = np.array([first[k] for k in ka.tolist() if k in first], dtype=int)
    return mask, idx


def is_row_in(a, b):
    return ismember(a, b, method='rows')[0]
