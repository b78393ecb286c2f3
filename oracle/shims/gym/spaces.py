import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype)
        self.high = np.asarray(high, dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
