"""Advantage Actor-Critic on the device (reference: tonic/torch/agents/a2c.py).

`step` = actor forward (csrc/mlp.cu) + Gaussian sample / log-prob
(csrc/heads.cu); `update` writes the transition into the HBM segment and, when
it is full, runs the whole update as a stream of kernels with one read-back of
the statistics at the end.
"""

import numpy as np
import torch

from ... import _lib, config, distributed, kernels, replays
from ...utils import logger
from .. import models, normalizers, updaters
from . import agent


def default_model():
    return models.ActorCritic(
        actor=models.Actor(
            encoder=models.ObservationEncoder(),
            torso=models.MLP((64, 64), torch.nn.Tanh),
            head=models.DetachedScaleGaussianPolicyHead()),
        critic=models.Critic(
            encoder=models.ObservationEncoder(),
            torso=models.MLP((64, 64), torch.nn.Tanh),
            head=models.ValueHead()),
        observation_normalizer=normalizers.MeanStd())


class A2C(agent.Agent):
    def __init__(self, model=None, replay=None, actor_updater=None, critic_updater=None):
        self.model = model or default_model()
        self.replay = replay or replays.Segment()
        self.actor_updater = actor_updater or updaters.StochasticPolicyGradient()
        self.critic_updater = critic_updater or updaters.VRegression()

    def initialize(self, observation_space, action_space, seed=None):
        super().initialize(seed=seed)
        self.model.initialize(observation_space, action_space)
        self.replay.initialize(seed)
        self.actor_updater.initialize(self.model)
        self.critic_updater.initialize(self.model)
        self.action_size = action_space.shape[0]
        self._noise_counter = 0
        self._workers = 0

    # -- acting -----------------------------------------------------------------
    def _buffers(self, workers):
        if workers != self._workers:
            dev = kernels.device()
            A = self.action_size
            self._pre = torch.empty(workers, A, dtype=torch.float32, device=dev)
            self._actions = torch.empty(workers, A, dtype=torch.float32, device=dev)
            self._log_probs = torch.empty(workers, dtype=torch.float32, device=dev)
            self._workers = workers

    def _sample(self, observations, actions, log_probs):
        """Normal(loc, scale).sample() and its summed log-prob (a2c.py:75-85)."""
        workers = observations.shape[0]
        self.model.actor.pre_activations(observations, out=self._pre[:workers])
        world, rank = distributed.world(), distributed.rank()
        eps, counter = None, self._noise_counter + rank * workers
        if config.noise == 'host':
            # same draw as torch.distributions.Normal.sample() from the global CPU
            # generator; with several ranks every rank draws the global block and
            # keeps its workers' rows (identical to the single-process stream)
            eps = torch.randn(workers * world, self.action_size)
            eps = eps[rank * workers:(rank + 1) * workers].to(observations.device)
        else:
            self._noise_counter += workers * world
        kernels.gauss_sample(self._pre[:workers], self.model.actor.network.extra('log_scale'),
                             actions, log_probs, eps=eps, seed=self.seed or 0, counter=counter)

    def step(self, observations, steps):
        host = not (isinstance(observations, torch.Tensor) and observations.is_cuda)
        observations = kernels.to_device(observations)
        self._buffers(observations.shape[0])
        self._sample(observations, self._actions, self._log_probs)
        # values kept for the next update (a2c.py:48-50)
        self.last_observations = observations.clone() if not host else observations
        self.last_actions = self._actions
        self.last_log_probs = self._log_probs
        return kernels.to_host(self._actions) if host else self._actions

    def test_step(self, observations, steps):
        host = not (isinstance(observations, torch.Tensor) and observations.is_cuda)
        observations = kernels.to_device(observations)
        dev = observations.device
        actions = torch.empty(observations.shape[0], self.action_size, device=dev)
        log_probs = torch.empty(observations.shape[0], device=dev)
        self._buffers(max(self._workers, observations.shape[0]))
        self._sample(observations, actions, log_probs)      # a2c.py:87-90: stochastic
        return kernels.to_host(actions) if host else actions

    # -- fused collection (device environments) ------------------------------------
    def rollout(self, environment, vector_steps, steps=0, action_stats=None):
        """Runs up to `vector_steps` steps of agent.step -> environment.step ->
        agent.update (trainer.py:44-50) without leaving the device: the acting
        observations of step t live in row t of the segment, the environment
        kernel writes next_observations / rewards / resets / terminations of step t
        and the acting observations of step t+1 straight into the segment rows.
        Stops when the segment is full (and then runs the update).  Returns the
        number of vector steps done."""
        seg, env = self.replay, environment
        N, A = env.workers, self.action_size
        O = env.observation_space.shape[0]
        if seg.buffers is None:
            seg.allocate(observations=(N, O), actions=(N, A), next_observations=(N, O),
                         rewards=(N,), resets=(N,), terminations=(N,), log_probs=(N,))
        self._buffers(N)
        b, T = seg.buffers, seg.max_size
        normalizer = self.model.observation_normalizer
        done = 0
        while done < vector_steps and seg.index < T:
            t = seg.index
            if t == 0:       # first acting observations come from the environment
                b['observations'][0].copy_(env.observations)
            obs = b['observations'][t]
            self._sample(obs, b['actions'][t], b['log_probs'][t])
            if normalizer:
                normalizer.record(obs)
            target = b['observations'][t + 1] if t + 1 < T else env.observations
            env.step_into(b['actions'][t], target, b['next_observations'][t], b['rewards'][t],
                          b['resets'][t], b['terminations'][t])
            seg.advance()
            done += 1
        if action_stats is not None and done:
            action_stats.add(b['actions'][seg.index - done:seg.index], items=done)
        if seg.ready():
            self._update()
        return done

    # -- learning ---------------------------------------------------------------
    def update(self, observations, rewards, resets, terminations, steps):
        self.replay.store(
            observations=self.last_observations, actions=self.last_actions,
            next_observations=observations, rewards=rewards, resets=resets,
            terminations=terminations, log_probs=self.last_log_probs)
        if self.model.observation_normalizer:
            self.model.observation_normalizer.record(self.last_observations)
        if self.replay.ready():
            self._update()

    def _evaluate(self):
        """V(s), V(s') for the whole segment, lambda-returns (a2c.py:92-105)."""
        flat = self.replay.get_full('observations', 'next_observations')
        total = flat['observations'].shape[0]
        dev = flat['observations'].device
        values = torch.empty(total, 1, dtype=torch.float32, device=dev)
        next_values = torch.empty(total, 1, dtype=torch.float32, device=dev)
        self.model.critic.values(flat['observations'], out=values)
        self.model.critic.values(flat['next_observations'], out=next_values)
        self.replay.compute_returns(values, next_values)

    def _stats(self, n):
        return torch.zeros(n, 2, _lib.STAT_COUNT, dtype=torch.float64, device=kernels.device())

    def _update(self):
        self._evaluate()
        flat = self.replay.get_full('observations', 'actions', 'advantages', 'log_probs',
                                    'returns')
        total = flat['observations'].shape[0]
        batches = list(self.replay.index_batches())
        stats = self._stats(len(batches) + 1)
        # one policy-gradient step on the full batch (a2c.py:107-114)
        self.actor_updater.launch(flat['observations'], flat['actions'], flat['advantages'],
                                  flat['log_probs'], None, total, stats[0, 0],
                                  rows_global=total * distributed.world())
        # several value-regression steps (a2c.py:116-121)
        for j, (idx, rows, rows_global) in enumerate(batches):
            self.critic_updater.launch(flat['observations'], flat['returns'], idx, rows,
                                       stats[j + 1, 1], rows_global=rows_global)
        host = kernels.to_host(stats)
        for k, v in self.actor_updater.infos(host[0, 0]).items():
            logger.store('actor/' + k, v)
        for j in range(len(batches)):
            for k, v in self.critic_updater.infos(host[j + 1, 1]).items():
                logger.store('critic/' + k, v)
        if self.model.observation_normalizer:
            self.model.observation_normalizer.update()
