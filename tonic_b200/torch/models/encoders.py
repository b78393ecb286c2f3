"""Input encoders (reference: tonic/torch/models/encoders.py:4-31).  They only
describe how the first-layer input is assembled -- [normalise(observations) |
actions] -- the assembly itself is fused into the first layer of
csrc/mlp.cu::mlp_forward_kernel."""

import torch


class ObservationEncoder(torch.nn.Module):
    def initialize(self, observation_space, action_space=None, observation_normalizer=None):
        self.observation_normalizer = observation_normalizer
        self.action_size = 0
        return observation_space.shape[0]


class ObservationActionEncoder(torch.nn.Module):
    def initialize(self, observation_space, action_space, observation_normalizer=None):
        self.observation_normalizer = observation_normalizer
        self.action_size = action_space.shape[0]
        return observation_space.shape[0] + self.action_size
