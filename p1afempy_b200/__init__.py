"""p1afempy_b200 -- B200-native P1 assembly hot path, drop-in for the
vectorised builders of leuraph/p1afempy (solvers / indicators / mesh helpers).

    from p1afempy_b200 import solvers, indicators, mesh     # numpy in, scipy out
    from p1afempy_b200.device import DeviceMesh              # device-resident API

Importing the package does not need a GPU; calling anything that computes
does, and raises if libp1b200.so is not built (no CPU fallback).
"""
__version__ = '0.1.0'
