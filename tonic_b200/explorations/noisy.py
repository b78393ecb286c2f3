"""Exploration for the deterministic / soft off-policy agents (reference:
tonic/explorations/noisy.py:6-50).

`policy(observations, noise)` is the agent's device policy; this object only
decides WHICH actions are taken (uniform warm-up for `steps <= start_steps`,
then the policy, optionally with Normal(0, scale) noise) and owns the noise
stream.  With `config.noise == 'host'` the numbers come from the native
numpy-compatible `RandomState(seed)` (bit-identical to the reference's stream:
float64 uniform warm-up actions, float64 normal noise added to float32 actions);
with 'device' they are drawn by Philox inside the action kernel.
The OU process (noisy.py:53-88) is a "next" row.
"""

import numpy as np
import torch

from .. import config, distributed, kernels
from ..utils.random_state import RandomState


class NoActionNoise:
    scale = 0.0

    def __init__(self, start_steps=20000):
        self.start_steps = start_steps

    def initialize(self, policy, action_space, seed=None):
        self.policy = policy
        self.action_size = action_space.shape[0]
        self.np_random = RandomState(seed)
        self.seed = seed or 0
        self._counter = 0

    def warmup_actions(self, workers):
        """Uniform(-1, 1) actions (noisy.py:20-21,44-46) as a device float32 tensor."""
        if config.noise == 'host':
            return kernels.to_device(self._own(
                self.np_random.uniform(-1, 1, (workers * distributed.world(), self.action_size)),
                workers))
        out = torch.empty(workers, self.action_size, dtype=torch.float32, device=kernels.device())
        kernels.tanh_action(None, out, mode=2, seed=self.seed ^ 0x5eed, counter=self._counter)
        self._counter += workers
        return out

    def noise(self, workers):
        return None

    @staticmethod
    def _own(block, workers):
        """Rows of the global (single-process) draw that belong to this rank."""
        rank = distributed.rank()
        return block[rank * workers:(rank + 1) * workers]

    def __call__(self, observations, steps):
        if steps > self.start_steps:
            return self.policy(observations, self.noise(len(observations)))
        return self.warmup_actions(len(observations))

    def update(self, resets):
        pass


class NormalActionNoise(NoActionNoise):
    def __init__(self, scale=0.1, start_steps=20000):
        self.scale = scale
        self.start_steps = start_steps

    def noise(self, workers):
        """float64 standard normals from the numpy-compatible stream, or None to
        let the kernel draw Philox noise."""
        if config.noise == 'host':
            return kernels.to_device(self._own(
                self.np_random.normal((workers * distributed.world(), self.action_size)),
                workers), dtype=torch.float64)
        return None
