import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (B200)')


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) on machines without a CUDA device or
    without the built library; on a GPU box a missing library still fails loudly inside
    the product (`tonic_b200/_lib.py`)."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device (B200)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
        return cache[name]
    return load


@pytest.fixture(autouse=True)
def _restore_logger():
    """tests/product.py::build replaces `logger.store` with a recorder; put the module back
    the way it was so later tests (the train CLI) log into their own logger."""
    try:
        from tonic_b200.utils import logger
    except Exception:       # package not importable in this environment: nothing to restore
        yield
        return
    saved = {k: getattr(logger, k) for k in ('store', 'store_aggregate', 'dump', 'current_logger')}
    yield
    for k, v in saved.items():
        setattr(logger, k, v)
