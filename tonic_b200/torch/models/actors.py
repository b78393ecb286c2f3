"""Policy networks (reference: tonic/torch/models/actors.py).

Heads only hold parameters and their hyper-parameters; the head arithmetic
(tanh / softplus / clamp / sampling / log-probabilities) is done by
csrc/heads.cu on the pre-activations produced by csrc/mlp.cu.
"""

import torch

from ... import kernels
from . import network

FLOAT_EPSILON = 1e-8


class DetachedScaleGaussianPolicyHead(torch.nn.Module):
    """Reference: actors.py:37-66 -- loc = act(Linear), state-independent
    scale = clamp(softplus(log_scale) + 1e-8, scale_min, scale_max)."""

    kind = 'detached_gaussian'

    def __init__(self, loc_activation=torch.nn.Tanh, loc_fn=None, log_scale_init=0.,
                 scale_min=1e-4, scale_max=1., distribution=torch.distributions.normal.Normal):
        super().__init__()
        if loc_activation is not torch.nn.Tanh or (scale_min, scale_max) != (1e-4, 1.) or \
                distribution is not torch.distributions.normal.Normal:
            raise NotImplementedError('only the reference defaults have a kernel')
        self.loc_activation, self.loc_fn = loc_activation, loc_fn
        self.log_scale_init = log_scale_init

    def initialize(self, input_size, action_size):
        self.loc_layer = torch.nn.Sequential(
            torch.nn.Linear(input_size, action_size), self.loc_activation())
        if self.loc_fn:
            self.loc_layer.apply(self.loc_fn)
        self.log_scale = torch.nn.Parameter(torch.as_tensor(
            [[self.log_scale_init] * action_size], dtype=torch.float32))

    def linears(self):
        return [self.loc_layer[0]]

    def extras(self):
        return [('log_scale', self.log_scale)]


class GaussianPolicyHead(torch.nn.Module):
    """Reference: actors.py:69-98 with the SAC configuration (loc Identity,
    scale = clamp(softplus(Linear), 1e-4, 1), tanh-squashed distribution)."""

    kind = 'squashed_gaussian'

    def __init__(self, loc_activation=torch.nn.Identity, loc_fn=None,
                 scale_activation=torch.nn.Softplus, scale_min=1e-4, scale_max=1,
                 scale_fn=None, distribution=None):
        super().__init__()
        if loc_activation is not torch.nn.Identity or scale_activation is not torch.nn.Softplus \
                or (scale_min, scale_max) != (1e-4, 1):
            raise NotImplementedError('only the SAC configuration has a kernel')
        self.loc_fn, self.scale_fn = loc_fn, scale_fn
        self.distribution = distribution or SquashedMultivariateNormalDiag

    def initialize(self, input_size, action_size):
        self.loc_layer = torch.nn.Sequential(
            torch.nn.Linear(input_size, action_size), torch.nn.Identity())
        if self.loc_fn:
            self.loc_layer.apply(self.loc_fn)
        self.scale_layer = torch.nn.Sequential(
            torch.nn.Linear(input_size, action_size), torch.nn.Softplus())
        if self.scale_fn:
            self.scale_layer.apply(self.scale_fn)

    def linears(self):
        return [self.loc_layer[0], self.scale_layer[0]]

    def extras(self):
        return []


class SquashedMultivariateNormalDiag:
    """Marker for the tanh-squashed Gaussian (actors.py:7-34); the sampling and
    log-prob arithmetic is in csrc/heads.cu."""


class DeterministicPolicyHead(torch.nn.Module):
    """Reference: actors.py:101-115 -- action = tanh(Linear)."""

    kind = 'deterministic'

    def __init__(self, activation=torch.nn.Tanh, fn=None):
        super().__init__()
        if activation is not torch.nn.Tanh:
            raise NotImplementedError('only tanh actions have a kernel')
        self.activation, self.fn = activation, fn

    def initialize(self, input_size, action_size):
        self.action_layer = torch.nn.Sequential(
            torch.nn.Linear(input_size, action_size), self.activation())
        if self.fn is not None:
            self.action_layer.apply(self.fn)

    def linears(self):
        return [self.action_layer[0]]

    def extras(self):
        return []


class Actor(torch.nn.Module):
    """Reference: actors.py:118-137.

    `normalize_observations=False` reproduces the reference's behaviour: its
    `Actor.initialize` passes the normaliser positionally into the encoder's
    `action_space` slot (actors.py:128-129 vs encoders.py:5-8), so torch actors
    see RAW observations while critics see normalised ones.
    """

    def __init__(self, encoder, torso, head, normalize_observations=False):
        super().__init__()
        self.encoder, self.torso, self.head = encoder, torso, head
        self.normalize_observations = normalize_observations

    def initialize(self, observation_space, action_space, observation_normalizer=None):
        normalizer = observation_normalizer if self.normalize_observations else None
        size = self.encoder.initialize(observation_space, observation_normalizer=normalizer)
        size = self.torso.initialize(size)
        self.action_size = action_space.shape[0]
        self.head.initialize(size, self.action_size)
        self.network = network.BoundNetwork(self.torso, self.head.linears(), self.head.extras())

    # -- kernels ------------------------------------------------------------
    def input(self, observations, idx=None):
        norm = self.encoder.observation_normalizer
        return kernels.MlpInput(observations, None if norm is None else norm._mean.data,
                                None if norm is None else norm._std.data, idx=idx)

    def pre_activations(self, observations, out=None, idx=None, rows=None, save=False,
                        skip=None):
        rows = observations.shape[0] if rows is None else rows
        if out is None:
            out = network.scratch(rows, self.network.layout.n_out, observations)
        return self.network.mlp.forward(self.input(observations, idx), rows, out, save=save,
                                        skip=skip)

    def forward(self, observations):
        """Convenience (not on the hot path): distribution / actions as torch objects."""
        observations = kernels.to_device(observations)
        pre = self.pre_activations(observations)
        if self.head.kind == 'deterministic':
            return torch.tanh(pre)
        if self.head.kind == 'detached_gaussian':
            scale = torch.clamp(torch.nn.functional.softplus(self.head.log_scale)
                                + FLOAT_EPSILON, 1e-4, 1.).repeat(pre.shape[0], 1)
            return torch.distributions.normal.Normal(torch.tanh(pre), scale)
        A = self.action_size
        scale = torch.clamp(torch.nn.functional.softplus(pre[:, A:]), 1e-4, 1)
        return torch.distributions.normal.Normal(pre[:, :A], scale)
