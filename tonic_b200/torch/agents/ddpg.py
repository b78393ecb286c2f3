"""Deep Deterministic Policy Gradient on the device (reference:
tonic/torch/agents/ddpg.py:20-112)."""

import numpy as np
import torch

from ... import _lib, config, distributed, explorations, graphs, kernels, replays
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
        self._noise_base = None          # device-resident Philox position (fast path)
        self._sections = {}              # captured (warm, updating) vector-step graphs
        self._update_stats = None

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

    def _stats_block(self):
        n = self.replay.batch_iterations
        if self._update_stats is None or self._update_stats.shape[0] != n:
            self._update_stats = torch.zeros(n, 2, _lib.STAT_COUNT, dtype=torch.float64,
                                             device=kernels.device())
        return self._update_stats

    def _enqueue_update(self, batches, stats):
        """All kernels of one update (ddpg.py:86-112), no host synchronisation."""
        stats.zero_()
        obs = self.replay.flat('observations')
        for i, (idx, rows, rows_global, mine) in enumerate(batches):  # ddpg.py:105-112, td3.py:41-46
            self.critic_updater.launch(self.replay, idx, rows, stats[i, 0],
                                       rows_global=rows_global, mine=mine)
            if self._actor_turn(i):
                self.actor_updater.launch(obs, idx, rows, stats[i, 1], rows_global=rows_global,
                                          mine=mine)
                self.model.update_targets()
        if self.model.observation_normalizer:
            self.model.observation_normalizer.update()

    def _report(self, host):
        for i in range(host.shape[0]):
            for k, v in self.critic_updater.infos(host[i, 0]).items():
                logger.store('critic/' + k, v)
            if self._actor_turn(i):
                for k, v in self.actor_updater.infos(host[i, 1]).items():
                    logger.store('actor/' + k, v)

    def _update(self, steps):
        stats = self._stats_block()
        self._enqueue_update(list(self.replay.index_batches(steps)), stats)
        self._report(kernels.to_host(stats))

    # -- fused collection (device environments, fast mode) ----------------------------------
    def _graphable(self):
        """Device-resident RNG streams and ring state: a whole vector step (act -> environment
        -> store -> record [-> update]) can be replayed as a CUDA graph."""
        return (config.graphs and self._device_mode()
                and (distributed.world() == 1 or config.graphs_multi_gpu))

    def _device_mode(self):
        return (config.noise == 'device' and config.indices == 'device'
                and not getattr(self.exploration, 'additive', False))

    def can_rollout(self, environment):
        """The fused trainer path (utils/trainer.py) needs the device-resident streams."""
        return self._device_mode() and self.replay.return_steps == 1

    def _reset_noise_offsets(self):
        """Offsets inside one vector step (the device base advances between steps)."""
        self._noise_counter = 0
        self.exploration._counter = 0
        self.actor_updater._counter = self.critic_updater._counter = 0

    def _enqueue_step(self, env, warm, updating):
        """agent.step -> environment.step -> agent.update of trainer.py:44-50 for one vector
        step, entirely on the device: the transition goes from fixed staging buffers into the
        ring row the DEVICE ring state points to."""
        rep = self.replay
        _lib.call('tb_set_noise_base', _lib.ptr(self._noise_base))
        try:
            self._reset_noise_offsets()
            obs = self._staged_obs
            obs.copy_(env.observations)                      # a2c/ddpg keep a copy of the acting observations
            if warm:
                actions = self._policy(obs, self.exploration.noise(obs.shape[0]))
            else:
                actions = self.exploration.warmup_actions(obs.shape[0])
            self._staged_actions.copy_(actions)
            env.step_into(self._staged_actions, env.observations, env.next_observations, env.rewards,
                          env.resets, env.terminations)
            rep.store_device(observations=obs, actions=self._staged_actions,
                             next_observations=env.next_observations, rewards=env.rewards,
                             resets=env.resets, terminations=env.terminations)
            rep.advance_device()
            if self.model.observation_normalizer:
                self.model.observation_normalizer.record(obs)
            if updating:
                self._enqueue_update(list(rep.index_batches_device(self.seed)), self._stats_block())
            kernels.counter_add(self._noise_base, 1 << 24)
        finally:
            _lib.call('tb_set_noise_base', None)

    def rollout(self, environment, vector_steps, steps=0, action_stats=None):
        """Runs `vector_steps` iterations of the training loop's body (trainer.py:44-55) without
        leaving the device; every iteration that `replay.ready` is followed by the update.
        Fast mode only (device noise and indices); returns the number of vector steps done."""
        if not self._device_mode() or self.replay.return_steps > 1:
            raise NotImplementedError('rollout() needs config.noise == config.indices == "device", '
                                      'a non-OU exploration and return_steps == 1; use step/update')
        env, rep = environment, self.replay
        N, A, world = env.workers, self.action_size, distributed.world()
        O = env.observation_space.shape[0]
        if rep.buffers is None:
            rep.allocate(observations=(N, O), actions=(N, A), next_observations=(N, O),
                         rewards=(N,), resets=(N,), terminations=(N,))
        if self._noise_base is None:
            dev = kernels.device()
            self._noise_base = kernels.new_counter()
            self._staged_obs = torch.empty(N, O, dtype=torch.float32, device=dev)
            self._staged_actions = torch.empty(N, A, dtype=torch.float32, device=dev)
        if getattr(rep, '_ring_stale', True):
            rep.sync_ring()
            rep._ring_stale = False
        for done in range(vector_steps):
            now = steps + done * N * world
            updating = rep.ready(now)
            warm = now > self.exploration.start_steps
            key = (warm, updating, id(env))
            if self._graphable():
                if key not in self._sections:
                    self._sections[key] = graphs.CapturedSection(
                        lambda w=warm, u=updating: self._enqueue_step(env, w, u))
                self._sections[key]()
            else:
                self._enqueue_step(env, warm, updating)
            rep.advance(mirror_only=True)
            if action_stats is not None:
                action_stats.add(self._staged_actions)
            if updating:
                rep.last_steps = now
                self._report(kernels.to_host(self._stats_block()))
        return vector_steps
