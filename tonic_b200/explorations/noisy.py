"""Exploration for the deterministic / soft off-policy agents (reference:
tonic/explorations/noisy.py:6-50).

`policy(observations, noise)` is the agent's device policy; this object only
decides WHICH actions are taken (uniform warm-up for `steps <= start_steps`,
then the policy, optionally with Normal(0, scale) noise) and owns the noise
stream.  With `config.noise == 'host'` the numbers come from the native
numpy-compatible `RandomState(seed)` (bit-identical to the reference's stream:
float64 uniform warm-up actions, float64 normal noise added to float32 actions);
with 'device' they are drawn by Philox inside the action kernel.
The OU process (noisy.py:53-88) keeps its state on the host (numpy float32 arithmetic of the
reference, numpy-compatible normal stream) and hands the policy an additive float32 noise block.
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


class OrnsteinUhlenbeckActionNoise(NoActionNoise):
    """Reference: explorations/noisy.py:53-88.  `noise()` advances the process exactly like the
    reference (`noises -= theta * noises * dt; noises += scale * sqrt(dt) * clip(N(0,1), +-clip)`,
    float32 state updated in place with numpy's casting rules) and returns the rows of this rank
    as an ADDITIVE float32 block; `update(resets)` clears the state of finished episodes.  The
    normal draws always come from the host stream (the state is host-resident)."""

    additive = True

    def __init__(self, scale=0.1, clip=2, theta=.15, dt=1e-2, start_steps=20000):
        self.scale, self.clip, self.theta, self.dt = scale, clip, theta, dt
        self.start_steps = start_steps

    def initialize(self, policy, action_space, seed=None):
        super().initialize(policy, action_space, seed)
        self.noises = None

    def warmup_actions(self, workers):      # uniform(-1, 1) from the host stream (noisy.py:82-84)
        return kernels.to_device(self._own(
            self.np_random.uniform(-1, 1, (workers * distributed.world(), self.action_size)), workers))

    def noise(self, workers):
        if self.noises is None:
            self.noises = np.zeros((workers, self.action_size), np.float32)     # zeros_like(actions)
        draws = self._own(self.np_random.normal((workers * distributed.world(), self.action_size)),
                          workers)
        draws = np.clip(draws, -self.clip, self.clip)
        self.noises -= self.theta * self.noises * self.dt
        self.noises += self.scale * np.sqrt(self.dt) * draws
        return kernels.to_device(self.noises)

    def update(self, resets):
        if self.noises is not None:
            resets = kernels.to_host(resets) if isinstance(resets, torch.Tensor) else np.asarray(resets)
            self.noises *= (1. - resets.astype(np.float64))[:, None]
