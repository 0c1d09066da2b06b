import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')
    # the suite uses the reference's rule names offline on purpose (oracle and CUDA
    # path are fed the same table); test_cubature_stand_ins checks the warning itself
    config.addinivalue_line('filterwarnings',
                            'ignore::p1afempy_b200.cubature.StandInRuleWarning')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _bit_exact_callbacks():
    """The parity tests compare with the oracle bit for bit: user callbacks are
    evaluated by numpy on the host, as the reference does (the package default
    evaluates them on the device; tests of that mode set it themselves)."""
    from p1afempy_b200 import options
    saved = options.callbacks
    options.callbacks = 'host'
    yield
    options.callbacks = saved
