"""Running observation normaliser (reference:
tonic/torch/normalizers/mean_stds.py:5-74).  `record` and `update` are the
kernels of csrc/moments.cu; `forward` is fused into the first MLP layer
(csrc/mlp.cu), so this module only owns the statistics:

  _mean, _std   float32 parameters (same state_dict keys as the reference)
  running       [mean | mean_sq] float32 (the reference's numpy attributes)
  sums          float64 [sum x | sum x^2 | count] recorded since the last update
"""

import numpy as np
import torch

from ... import distributed, kernels


class MeanStd(torch.nn.Module):
    def __init__(self, mean=0, std=1, clip=None, shape=None):
        super().__init__()
        if clip is not None:
            raise NotImplementedError('clipping is not used by the in-scope agents')
        self._init_mean, self._init_std = mean, std
        self.eps = 1e-2
        if shape:
            self.initialize(shape)

    def initialize(self, shape):
        dev = kernels.device()
        size = int(np.prod(shape))
        mean = np.broadcast_to(np.asarray(self._init_mean, np.float32), shape).ravel()
        std = np.broadcast_to(np.asarray(self._init_std, np.float32), shape).ravel()
        self._mean = torch.nn.Parameter(torch.as_tensor(mean.copy(), device=dev),
                                        requires_grad=False)
        self._std = torch.nn.Parameter(torch.as_tensor(std.copy(), device=dev),
                                       requires_grad=False)
        self.running = torch.cat([self._mean.data, self._mean.data ** 2])
        self.sums = torch.zeros(2 * size + 1, dtype=torch.float64, device=dev)
        self.count_buffer = torch.zeros(1, dtype=torch.float64, device=dev)
        self.size = size

    # reference attribute names (host copies, for inspection)
    @property
    def mean(self):
        return self._mean.detach().cpu().numpy()

    @property
    def std(self):
        return self._std.detach().cpu().numpy()

    @property
    def count(self):
        return int(self.count_buffer.item())

    def forward(self, val):
        """(val - mean) / std -- convenience; the kernels fuse this into layer 1."""
        return (kernels.to_device(val) - self._mean) / self._std

    def unnormalize(self, val):
        return kernels.to_device(val) * self._std + self._mean

    def record(self, values):
        values = kernels.to_device(values)
        kernels.moments_record(values.view(-1, self.size), self.sums)

    def update(self):
        distributed.all_reduce(self.sums)       # global statistics when workers are sharded
        kernels.moments_update(self.sums, self.running, self.count_buffer, self._mean.data,
                               self._std.data, self.eps)
