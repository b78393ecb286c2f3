"""Value networks (reference: tonic/torch/models/critics.py:4-20,70-90)."""

import torch

from ... import kernels
from . import network


class ValueHead(torch.nn.Module):
    kind = 'value'

    def __init__(self, fn=None):
        super().__init__()
        self.fn = fn

    def initialize(self, input_size, return_normalizer=None):
        if return_normalizer is not None:
            raise NotImplementedError('return normalisers are outside the hot path '
                                      '(SURVEY.md section 2, row 17)')
        self.return_normalizer = None
        self.v_layer = torch.nn.Linear(input_size, 1)
        if self.fn:
            self.v_layer.apply(self.fn)

    def linears(self):
        return [self.v_layer]

    def extras(self):
        return []


class Critic(torch.nn.Module):
    def __init__(self, encoder, torso, head):
        super().__init__()
        self.encoder, self.torso, self.head = encoder, torso, head

    def initialize(self, observation_space, action_space, observation_normalizer=None,
                   return_normalizer=None):
        size = self.encoder.initialize(
            observation_space=observation_space, action_space=action_space,
            observation_normalizer=observation_normalizer)
        size = self.torso.initialize(size)
        self.head.initialize(size, return_normalizer)
        self.network = network.BoundNetwork(self.torso, self.head.linears(), self.head.extras())

    def input(self, observations, actions=None, idx=None, gather_actions=True):
        norm = self.encoder.observation_normalizer
        return kernels.MlpInput(observations, None if norm is None else norm._mean.data,
                                None if norm is None else norm._std.data, x2=actions,
                                gather2=gather_actions and idx is not None, idx=idx)

    def values(self, observations, actions=None, out=None, idx=None, rows=None,
               gather_actions=True, save=False, skip=None, params=None, packed=None, vloss=None):
        rows = observations.shape[0] if rows is None else rows
        if out is None:
            out = network.scratch(rows, 1, observations)
        inp = self.input(observations, actions, idx, gather_actions)
        return self.network.mlp.forward(inp, rows, out, save=save, skip=skip, params=params,
                                        packed=packed, vloss=vloss)

    def forward(self, observations, actions=None):
        observations = kernels.to_device(observations)
        if actions is not None:
            actions = kernels.to_device(actions)
        return self.values(observations, actions).squeeze(-1)
