"""Empty stand-in so ``from matplotlib import pyplot`` in the reference
(p1afempy/mesh.py:6-7) resolves offline; only show_mesh() would use it."""
