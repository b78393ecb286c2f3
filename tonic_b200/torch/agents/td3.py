"""Twin Delayed DDPG on the device (reference: tonic/torch/agents/td3.py:20-55)."""

import torch

from .. import models, normalizers, updaters
from . import ddpg


def default_model():
    return models.ActorTwinCriticWithTargets(
        actor=models.Actor(
            encoder=models.ObservationEncoder(),
            torso=models.MLP((256, 256), torch.nn.ReLU),
            head=models.DeterministicPolicyHead()),
        critic=models.Critic(
            encoder=models.ObservationActionEncoder(),
            torso=models.MLP((256, 256), torch.nn.ReLU),
            head=models.ValueHead()),
        observation_normalizer=normalizers.MeanStd())


class TD3(ddpg.DDPG):
    def __init__(self, model=None, replay=None, exploration=None, actor_updater=None,
                 critic_updater=None, delay_steps=2):
        model = model or default_model()
        critic_updater = critic_updater or updaters.TwinCriticDeterministicQLearning()
        super().__init__(model=model, replay=replay, exploration=exploration,
                         actor_updater=actor_updater, critic_updater=critic_updater)
        self.delay_steps = delay_steps
        self.model.critic = self.model.critic_1                # td3.py:36

    def _actor_turn(self, iteration):                          # td3.py:42
        return (iteration + 1) % self.delay_steps == 0
