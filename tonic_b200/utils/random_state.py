"""Native (C++) numpy-compatible legacy `RandomState` streams.

The reference draws minibatch permutations, replay indices and exploration
noise from `numpy.random.RandomState(seed)` (tonic/replays/segments.py:20,62;
tonic/replays/buffers.py:22,86; tonic/explorations/noisy.py:13,21,41).  This
wrapper reproduces those streams bit-for-bit from the native library, writing
straight into caller-provided (pinned) host buffers so they can be DMA'd to the
device without going through the Python interpreter per element.
"""

import ctypes

import numpy as np

from .. import _lib


class RandomState:
    def __init__(self, seed=None):
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        if not 0 <= int(seed) < 2 ** 32:
            raise ValueError('seed must be between 0 and 2**32 - 1')
        self._lib = _lib.load()
        self._state = self._lib.tb_rs_create(int(seed))

    def __del__(self):
        state, self._state = getattr(self, '_state', None), None
        if state:
            self._lib.tb_rs_destroy(state)

    def shuffle(self, x):
        """In-place shuffle of a contiguous 1-D int64 numpy array."""
        assert x.dtype == np.int64 and x.ndim == 1 and x.flags.c_contiguous
        self._lib.tb_rs_shuffle_i64(self._state, x.ctypes.data, x.size)

    def randint(self, high, size, out=None):
        if out is None:
            out = np.empty(size, np.int64)
        assert out.dtype == np.int64 and out.size == int(np.prod(size))
        self._lib.tb_rs_randint(self._state, int(high), out.ctypes.data, out.size)
        return out

    def uniform(self, low, high, size):
        out = np.empty(size, np.float64)
        self._lib.tb_rs_uniform(self._state, float(low), float(high), out.ctypes.data, out.size)
        return out

    def normal(self, size):
        out = np.empty(size, np.float64)
        self._lib.tb_rs_normal(self._state, out.ctypes.data, out.size)
        return out
