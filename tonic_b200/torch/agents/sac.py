"""Soft Actor-Critic on the device (reference: tonic/torch/agents/sac.py:22-51)."""

import torch

from ... import config, distributed, explorations, kernels
from .. import models, normalizers, updaters
from . import ddpg


def default_model():
    return models.ActorTwinCriticWithTargets(
        actor=models.Actor(
            encoder=models.ObservationEncoder(),
            torso=models.MLP((256, 256), torch.nn.ReLU),
            head=models.GaussianPolicyHead(
                loc_activation=torch.nn.Identity,
                distribution=models.SquashedMultivariateNormalDiag)),
        critic=models.Critic(
            encoder=models.ObservationActionEncoder(),
            torso=models.MLP((256, 256), torch.nn.ReLU),
            head=models.ValueHead()),
        observation_normalizer=normalizers.MeanStd())


class SAC(ddpg.DDPG):
    def __init__(self, model=None, replay=None, exploration=None, actor_updater=None,
                 critic_updater=None):
        model = model or default_model()
        exploration = exploration or explorations.NoActionNoise()
        actor_updater = actor_updater or updaters.TwinCriticSoftDeterministicPolicyGradient()
        critic_updater = critic_updater or updaters.TwinCriticSoftQLearning()
        super().__init__(model=model, replay=replay, exploration=exploration,
                         actor_updater=actor_updater, critic_updater=critic_updater)

    def _policy(self, observations, noise=None):               # sac.py:40-46
        pre = self.model.actor.pre_activations(observations)
        out = self._new_actions(observations)
        eps = None
        if config.noise == 'host':      # Normal.sample() from torch's global CPU generator
            workers, world, rank = observations.shape[0], distributed.world(), distributed.rank()
            eps = torch.randn(workers * world, self.action_size)
            eps = eps[rank * workers:(rank + 1) * workers].to(pre.device)
        kernels.squashed_sample(pre, out, eps=eps, seed=(self.seed or 0) ^ 0x5ac0,
                                counter=self._noise_counter)
        self._noise_counter += observations.shape[0]
        return out

    def _greedy_actions(self, observations):                   # sac.py:48-51
        pre = self.model.actor.pre_activations(observations)
        out = self._new_actions(observations)
        kernels.squashed_sample(pre, out, greedy=True)
        return out
