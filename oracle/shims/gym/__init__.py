"""Test-only import shim for the `gym` package (absent in this image).

The unmodified reference (`/root/reference/tonic`) imports `gym`, `gym.wrappers`
and `gym.spaces` at module import time (environments/builders.py:5,
environments/wrappers.py:3).  This shim provides just the names those modules
touch so the reference can be imported as the parity oracle when generating
golden vectors (oracle/make_golden.py).  It is NOT part of the product.
"""
from . import core, spaces, wrappers  # noqa: F401
from .core import Env, Wrapper, ActionWrapper  # noqa: F401


def make(*args, **kwargs):
    raise NotImplementedError('gym is not installed; only custom envs work')
