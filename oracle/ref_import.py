"""Import the REAL reference (leuraph/p1afempy) from /root/reference.

Test infrastructure only.  Works in the build container (where the reference
is mounted read-only); on the GPU box ``/root/reference`` does not exist and
``available()`` returns False, so nothing at run time depends on it.

Three third-party modules the reference imports are absent offline
(``ismember``, ``matplotlib``, ``triangle_cubature``); ``oracle/shims`` holds
stand-ins that are put on ``sys.path`` first.
"""
import importlib
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get('P1_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'p1afempy'))


def data_dir() -> str:
    return os.path.join(REFERENCE_ROOT, 'p1afempy', 'tests', 'data')


def load():
    """Returns the imported reference package (``p1afempy``)."""
    if not available():
        raise RuntimeError('reference not present at ' + REFERENCE_ROOT)
    for path in (_REPO, REFERENCE_ROOT):
        if path not in sys.path:
            sys.path.append(path)
    for name in ('ismember', 'matplotlib', 'triangle_cubature'):
        try:
            importlib.import_module(name)
        except ImportError:
            if _SHIMS not in sys.path:
                sys.path.insert(0, _SHIMS)
            importlib.import_module(name)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', SyntaxWarning)
        ref = importlib.import_module('p1afempy')
        importlib.import_module('p1afempy.refinement')
        importlib.import_module('p1afempy.io_helpers')
    return ref
