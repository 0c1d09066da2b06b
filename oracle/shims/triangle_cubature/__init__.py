"""Stand-in for the third-party ``triangle-cubature`` package (absent offline,
unpinned in the reference's pyproject.toml:20).

The tables come from ``p1afempy_b200.cubature`` so that the reference (when
imported for oracle validation), the oracle and the CUDA path are always fed
the SAME numbers.  MIDPOINT is pinned by the reference's tests
(test_solver.py:88-129); DAYTAYLOR is a documented degree-6 stand-in ("parity
unpinned").
"""
