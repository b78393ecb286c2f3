"""Deep Deterministic Policy Gradient on the device (reference:
tonic/torch/agents/ddpg.py:20-112)."""

import numpy as np
import torch

from ... import _lib, explorations, kernels, replays
from ...utils import logger
from .. import models, normalizers, updaters
from . import agent


def default_model():
    return models.ActorCriticWithTargets(
        actor=models.Actor(
            encoder=models.ObservationEncoder(),
            torso=models.MLP((256, 256), torch.nn.ReLU),
            head=models.DeterministicPolicyHead()),
        critic=models.Critic(
            encoder=models.ObservationActionEncoder(),
            torso=models.MLP((256, 256), torch.nn.ReLU),
            head=models.ValueHead()),
        observation_normalizer=normalizers.MeanStd())


class DDPG(agent.Agent):
    def __init__(self, model=None, replay=None, exploration=None, actor_updater=None,
                 critic_updater=None):
        self.model = model or default_model()
        self.replay = replay or replays.Buffer()
        self.exploration = exploration or explorations.NormalActionNoise()
        self.actor_updater = actor_updater or updaters.DeterministicPolicyGradient()
        self.critic_updater = critic_updater or updaters.DeterministicQLearning()

    def initialize(self, observation_space, action_space, seed=None):
        super().initialize(seed=seed)
        self.model.initialize(observation_space, action_space)
        self.replay.initialize(seed)
        self.exploration.initialize(self._policy, action_space, seed)
        self.actor_updater.initialize(self.model)
        self.critic_updater.initialize(self.model)
        self.actor_updater.seed = self.critic_updater.seed = seed or 0
        self.action_size = action_space.shape[0]
        self._noise_counter = 0

    # -- acting -----------------------------------------------------------------
    def _new_actions(self, observations):
        return torch.empty(observations.shape[0], self.action_size, dtype=torch.float32,
                           device=observations.device)

    def _greedy_actions(self, observations):                  # ddpg.py:78-81
        pre = self.model.actor.pre_activations(observations)
        out = self._new_actions(observations)
        kernels.tanh_action(pre, out, mode=0)
        return out

    def _policy(self, observations, noise=None):
        """Greedy actions plus exploration noise (ddpg.py:83-84 + noisy.py:39-43):
        `noise` = float64 standard normals from the numpy stream, or None for
        in-kernel Philox noise."""
        scale = self.exploration.scale
        if not scale:
            return self._greedy_actions(observations)
        pre = self.model.actor.pre_activations(observations)
        out = self._new_actions(observations)
        if getattr(self.exploration, 'additive', False):
            # OU process: float32 noise state added as is (noisy.py:78-80)
            kernels.tanh_action(pre, out, mode=1, noise32=noise, noise_scale=1.0)
            return out
        kernels.tanh_action(pre, out, mode=1, noise64=noise, seed=(self.seed or 0) ^ 0xdd9,
                            counter=self._noise_counter, noise_scale=scale)
        self._noise_counter += observations.shape[0]
        return out

    def step(self, observations, steps):
        host = not (isinstance(observations, torch.Tensor) and observations.is_cuda)
        observations = kernels.to_device(observations)
        actions = self.exploration(observations, steps)
        self.last_observations = observations if host else observations.clone()
        self.last_actions = actions
        return kernels.to_host(actions) if host else actions

    def test_step(self, observations, steps):
        host = not (isinstance(observations, torch.Tensor) and observations.is_cuda)
        actions = self._greedy_actions(kernels.to_device(observations))
        return kernels.to_host(actions) if host else actions

    # -- learning ---------------------------------------------------------------
    def update(self, observations, rewards, resets, terminations, steps):
        self.replay.store(
            observations=self.last_observations, actions=self.last_actions,
            next_observations=observations, rewards=rewards, resets=resets,
            terminations=terminations)
        if self.model.observation_normalizer:
            self.model.observation_normalizer.record(self.last_observations)
        if self.replay.ready(steps):
            self._update(steps)
        self.exploration.update(resets)

    def _actor_turn(self, iteration):
        return True

    def _update(self, steps):
        batches = list(self.replay.index_batches(steps))
        stats = torch.zeros(len(batches), 2, _lib.STAT_COUNT, dtype=torch.float64,
                            device=kernels.device())
        obs = self.replay.flat('observations')
        for i, (idx, rows, rows_global, mine) in enumerate(batches):  # ddpg.py:105-112, td3.py:41-46
            self.critic_updater.launch(self.replay, idx, rows, stats[i, 0],
                                       rows_global=rows_global, mine=mine)
            if self._actor_turn(i):
                self.actor_updater.launch(obs, idx, rows, stats[i, 1], rows_global=rows_global,
                                          mine=mine)
                self.model.update_targets()
        host = kernels.to_host(stats)
        for i in range(len(batches)):
            for k, v in self.critic_updater.infos(host[i, 0]).items():
                logger.store('critic/' + k, v)
            if self._actor_turn(i):
                for k, v in self.actor_updater.infos(host[i, 1]).items():
                    logger.store('actor/' + k, v)
        if self.model.observation_normalizer:
            self.model.observation_normalizer.update()
