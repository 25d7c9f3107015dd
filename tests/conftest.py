import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _cuda_devices():
    try:
        from arpeggio_b200 import _lib
        return _lib.lib().arp_device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device (the build container) a plain `pytest` run skips the GPU tests instead of erroring in
    the engine fixture; `-m gpu` on the B200 box runs them."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason='no CUDA device: GPU parity tests run on the B200 box (pytest -m gpu)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def engine():
    """One ContactEngine on cuda:0 for the whole GPU session."""
    from arpeggio_b200.engine import ContactEngine
    eng = ContactEngine(device=0)
    yield eng
    eng.close()
