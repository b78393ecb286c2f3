"""Model containers (reference: tonic/torch/models/actor_critics.py:8-130)."""

import copy

import torch


class ActorCritic(torch.nn.Module):
    def __init__(self, actor, critic, observation_normalizer=None, return_normalizer=None):
        super().__init__()
        self.actor, self.critic = actor, critic
        self.observation_normalizer = observation_normalizer
        self.return_normalizer = return_normalizer

    def initialize(self, observation_space, action_space):
        if self.observation_normalizer:
            self.observation_normalizer.initialize(observation_space.shape)
        self.actor.initialize(observation_space, action_space, self.observation_normalizer)
        self.critic.initialize(observation_space, action_space, self.observation_normalizer,
                               self.return_normalizer)

    def networks(self):
        return [self.actor.network, self.critic.network]

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, *args, **kwargs):
        out = super().load_state_dict(state_dict, *args, **kwargs)
        for net in self.networks():
            net.refresh()
        return out


class _WithTargets(ActorCritic):
    """Shared machinery of the target-network containers: creation order and
    target assignment follow actor_critics.py:51-66 / 102-124, the soft update
    (:68-72, :126-130) is the fused kernel csrc/optim.cu::soft_update_kernel."""

    target_coeff = 0.005

    def _pairs(self):
        raise NotImplementedError

    def _finish(self):
        for online, target in self._pairs():
            for p in target.parameters():
                p.requires_grad = False
        self.assign_targets()

    def assign_targets(self):
        for online, target in self._pairs():
            target.network.copy_from(online.network)

    def update_targets(self):
        for online, target in self._pairs():
            target.network.soft_update_from(online.network, self.target_coeff)


class ActorCriticWithTargets(_WithTargets):
    def __init__(self, actor, critic, observation_normalizer=None, return_normalizer=None,
                 target_coeff=0.005):
        super().__init__(actor, critic, observation_normalizer, return_normalizer)
        self.target_actor = copy.deepcopy(actor)
        self.target_critic = copy.deepcopy(critic)
        self.target_coeff = target_coeff

    def initialize(self, observation_space, action_space):
        super().initialize(observation_space, action_space)
        self.target_actor.initialize(observation_space, action_space,
                                     self.observation_normalizer)
        self.target_critic.initialize(observation_space, action_space,
                                      self.observation_normalizer, self.return_normalizer)
        self._finish()

    def _pairs(self):
        return [(self.actor, self.target_actor), (self.critic, self.target_critic)]

    def networks(self):
        return [m.network for m in (self.actor, self.critic, self.target_actor,
                                    self.target_critic)]


class ActorTwinCriticWithTargets(_WithTargets):
    def __init__(self, actor, critic, observation_normalizer=None, return_normalizer=None,
                 target_coeff=0.005):
        torch.nn.Module.__init__(self)
        self.actor = actor
        self.critic_1 = critic
        self.critic_2 = copy.deepcopy(critic)
        self.target_actor = copy.deepcopy(actor)
        self.target_critic_1 = copy.deepcopy(critic)
        self.target_critic_2 = copy.deepcopy(critic)
        self.observation_normalizer = observation_normalizer
        self.return_normalizer = return_normalizer
        self.target_coeff = target_coeff

    def initialize(self, observation_space, action_space):
        if self.observation_normalizer:
            self.observation_normalizer.initialize(observation_space.shape)
        norm, ret = self.observation_normalizer, self.return_normalizer
        self.actor.initialize(observation_space, action_space, norm)
        self.critic_1.initialize(observation_space, action_space, norm, ret)
        self.critic_2.initialize(observation_space, action_space, norm, ret)
        self.target_actor.initialize(observation_space, action_space, norm)
        self.target_critic_1.initialize(observation_space, action_space, norm, ret)
        self.target_critic_2.initialize(observation_space, action_space, norm, ret)
        self._finish()

    def _pairs(self):
        return [(self.actor, self.target_actor), (self.critic_1, self.target_critic_1),
                (self.critic_2, self.target_critic_2)]

    def networks(self):
        return [m.network for m in (self.actor, self.critic_1, self.critic_2,
                                    self.target_actor, self.target_critic_1,
                                    self.target_critic_2)]
