"""Advantage Actor-Critic on the device (reference: tonic/torch/agents/a2c.py).

`step` = actor forward (csrc/mlp.cu) + Gaussian sample / log-prob
(csrc/heads.cu); `update` writes the transition into the HBM segment and, when
it is full, runs the whole update as a stream of kernels with one read-back of
the statistics at the end.
"""

import numpy as np
import torch

from ... import _lib, config, distributed, graphs, kernels, replays
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
        self._noise_counter = None      # device-resident Philox stream position
        self._workers = 0
        self._rollout_graph = self._update_graph = None
        self._update_buffers = None

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
        eps = None
        if self._noise_counter is None:
            self._noise_counter = kernels.new_counter()
        if config.noise == 'host':
            # same draw as torch.distributions.Normal.sample() from the global CPU
            # generator; with several ranks every rank draws the global block and
            # keeps its workers' rows (identical to the single-process stream)
            eps = torch.randn(workers * world, self.action_size)
            eps = eps[rank * workers:(rank + 1) * workers].to(observations.device)
        kernels.gauss_sample(self._pre[:workers], self.model.actor.network.extra('log_scale'),
                             actions, log_probs, eps=eps, seed=self.seed or 0,
                             counter=rank * workers, device_counter=self._noise_counter)
        if eps is None:     # same streams whatever the number of ranks
            kernels.counter_add(self._noise_counter, workers * world)

    def _host_fast_path(self):
        """Reference protocol with numpy arrays in the product configuration (device noise):
        each protocol call is one captured graph over fixed staging buffers (kernels.HostBridge)."""
        return config.graphs and config.noise == 'device' and distributed.world() == 1

    def _step_host(self, observations):
        """agent.step(numpy) -> numpy: pinned copy-in, H2D, actor forward, sample, D2H: one graph."""
        if getattr(self, '_bridge', None) is None:
            self._bridge, self._host_sections = kernels.HostBridge(), {}
        N = observations.shape[0]
        self._buffers(N)
        pin_obs, dev_obs = self._bridge.load('step_obs', observations)
        key = ('step', N)
        if key not in self._host_sections:
            last = self._bridge.buffers('last_obs', observations.shape)[1]

            def body():
                dev_obs.copy_(pin_obs, non_blocking=True)
                self._sample(dev_obs, self._actions[:N], self._log_probs[:N])
                last.copy_(dev_obs)                      # a2c.py:48: the agent keeps a copy
            # the actions come back through rotating pinned slots and are returned as views (no
            # host copy): valid until the third next call, like the environment's results
            slots = [torch.empty(N, self.action_size, dtype=torch.float32).pin_memory() for _ in range(3)]
            self._host_sections[key] = (graphs.CapturedSection(body), last, slots,
                                        [slot.numpy() for slot in slots], [0])
        section, last, slots, views, cursor = self._host_sections[key]
        section()
        cursor[0] = (cursor[0] + 1) % len(slots)
        slots[cursor[0]].copy_(self._actions[:N], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        kernels.transfers['d2h'] += slots[0].numel() * 4
        self.last_observations, self.last_actions, self.last_log_probs = last, self._actions[:N], self._log_probs[:N]
        return views[cursor[0]]

    def _update_host(self, observations, rewards, resets, terminations):
        """agent.update(numpy ...): pinned copy-in, 4 H2D copies, segment store at the device row
        index, normaliser record: one graph, no synchronisation."""
        seg, b = self.replay, self._bridge
        N = rewards.shape[0]
        if seg.buffers is None:
            O = observations.shape[1]
            seg.allocate(observations=(N, O), actions=(N, self.action_size), next_observations=(N, O),
                         rewards=(N,), resets=(N,), terminations=(N,), log_probs=(N,))
        seg.prepare_device_store()
        staged = [b.load('upd_next_obs', observations), b.load('upd_rewards', rewards),
                  b.load('upd_resets', resets), b.load('upd_terminations', terminations)]
        key = ('update', N)
        if key not in self._host_sections:
            last = self._bridge.buffers('last_obs', observations.shape)[1]
            normalizer = self.model.observation_normalizer

            def body():
                for pinned, dev in staged:
                    dev.copy_(pinned, non_blocking=True)
                seg.store_device(observations=last, actions=self._actions[:N],
                                 next_observations=staged[0][1], rewards=staged[1][1],
                                 resets=staged[2][1], terminations=staged[3][1],
                                 log_probs=self._log_probs[:N])
                if normalizer:
                    normalizer.record(last)
            self._host_sections[key] = (graphs.CapturedSection(body), None)
        self._host_sections[key][0]()
        seg.note_device_store()

    def step(self, observations, steps):
        host = not (isinstance(observations, torch.Tensor) and observations.is_cuda)
        if host and self._host_fast_path():
            return self._step_host(np.asarray(observations, np.float32))
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

        world, rank = distributed.world(), distributed.rank()
        fused_step = self._fused_step(env)

        def one_step(t, offset=0):
            """`offset`: vector steps since the last advance of the device noise counter (a whole
            captured segment advances it once, at the end)."""
            if t == 0:       # first acting observations come from the environment
                b['observations'][0].copy_(env.observations)
            obs = b['observations'][t]
            target = b['observations'][t + 1] if t + 1 < T else env.observations
            if fused_step:
                # actor forward, then ONE kernel: sample + log-prob, normaliser record, env step
                self.model.actor.pre_activations(obs, out=self._pre[:N])
                if self._noise_counter is None:
                    self._noise_counter = kernels.new_counter()
                kernels.act_env_step(
                    env.struct, self._pre[:N], self.model.actor.network.extra('log_scale'),
                    self.seed or 0, rank * N + offset * N * world, self._noise_counter,
                    b['actions'][t], b['log_probs'][t], normalizer.sums if normalizer else None,
                    target, b['next_observations'][t], b['rewards'][t], b['resets'][t],
                    b['terminations'][t])
                return
            self._sample(obs, b['actions'][t], b['log_probs'][t])
            if normalizer:
                normalizer.record(obs)
            env.step_into(b['actions'][t], target, b['next_observations'][t], b['rewards'][t],
                          b['resets'][t], b['terminations'][t])

        def whole_segment():
            for t in range(T):
                one_step(t, offset=t)
            if fused_step:
                kernels.counter_add(self._noise_counter, T * N * world)

        done = 0
        if self._fusable(env) and seg.index == 0 and vector_steps >= T:
            # whole segment in ONE persistent kernel (csrc/mlp.cu::rollout_kernel)
            if self._noise_counter is None:
                self._noise_counter = kernels.new_counter()
            actor = self.model.actor
            norm = actor.encoder.observation_normalizer
            world, rank = distributed.world(), distributed.rank()
            kernels.rollout_fused(
                env.struct, actor.network.mlp, actor.network.extra('log_scale'),
                None if norm is None else norm._mean.data,
                None if norm is None else norm._std.data, T, b, env.observations,
                normalizer.sums if normalizer else None, self.seed or 0, rank * N, N * world,
                self._noise_counter)
            kernels.counter_add(self._noise_counter, T * N * world)
            seg.index = T
            done = T
        elif self._graphable() and seg.index == 0 and vector_steps >= T:
            # whole segment in one CUDA graph (tonic_b200/graphs.py)
            if self._rollout_graph is None:
                self._rollout_graph = graphs.CapturedSection(whole_segment)
            self._rollout_graph()
            seg.index = T
            done = T
        while done < vector_steps and seg.index < T:
            one_step(seg.index)
            if fused_step:
                kernels.counter_add(self._noise_counter, N * world)
            seg.advance()
            done += 1
        if action_stats is not None and done:
            action_stats.add(b['actions'][seg.index - done:seg.index], items=done)
        if seg.ready():
            self._update()
        return done

    def _fused_step(self, env):
        """Device noise on the synthetic task: sampling, the normaliser record and the environment
        step of one vector step run as one kernel (csrc/env_step.cu::act_env_step_kernel)."""
        return (config.noise == 'device' and config.fused_step and hasattr(env, 'struct')
                and not getattr(env, 'time_feature', False)
                and not getattr(env.spec, 'task_id', 0)
                and env.struct.obs_dim <= 64 and self.action_size <= 64
                and hasattr(self.model.actor, 'pre_activations'))

    def _fusable(self, env):
        """The fused rollout kernel covers the detached-scale Gaussian actor on
        device noise with small observation / action vectors."""
        actor = self.model.actor
        shape = actor.network.layout.shape
        if config.fused_rollout == 'auto' and self._graphable() and actor.network.mlp.passes():
            return False        # graph replay of the tensor-core per-step chain is faster
        return (bool(config.fused_rollout) and config.noise == 'device'
                and getattr(actor.head, 'kind', None) == 'detached_gaussian'
                and hasattr(env, 'struct') and not getattr(env, 'time_feature', False)
                and not getattr(env.spec, 'task_id', 0)      # the kernel steps SynthControl only
                and shape.d_in <= min(64, shape.hidden)
                and shape.n_out <= 16 and shape.hidden in (64, 128, 256))

    def _graphable(self):
        """Static shapes + device-resident RNG: the section can be replayed as a CUDA
        graph.  With several ranks the NCCL all-reduces are captured too
        (config.graphs_multi_gpu)."""
        return (config.graphs and config.noise == 'device' and config.indices == 'device'
                and (distributed.world() == 1 or config.graphs_multi_gpu))

    # -- learning ---------------------------------------------------------------
    def update(self, observations, rewards, resets, terminations, steps):
        if isinstance(observations, np.ndarray) and self._host_fast_path() \
                and getattr(self, '_bridge', None) is not None \
                and isinstance(self.last_observations, torch.Tensor) \
                and ('step', observations.shape[0]) in self._host_sections \
                and self.last_observations is self._host_sections[('step', observations.shape[0])][1]:
            self._update_host(observations, np.asarray(rewards), np.asarray(resets), np.asarray(terminations))
            if self.replay.ready():
                self._update()
            return
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
        if getattr(self, '_value_buffers', None) is None or self._value_buffers[0].shape[0] != total:
            self._value_buffers = (torch.empty(total, 1, dtype=torch.float32, device=dev),
                                   torch.empty(total, 1, dtype=torch.float32, device=dev))
        values, next_values = self._value_buffers
        self.model.critic.values(flat['observations'], out=values)
        self.model.critic.values(flat['next_observations'], out=next_values)
        self.replay.compute_returns(values, next_values)

    def _buffers_for_update(self, n_batches):
        """Persistent device blocks (graph-safe): statistics [n, 2, TB_STAT_COUNT] and
        the KL early-stop flag."""
        if self._update_buffers is None or self._update_buffers[0].shape[0] != n_batches:
            dev = kernels.device()
            self._update_buffers = (
                torch.zeros(n_batches, 2, _lib.STAT_COUNT, dtype=torch.float64, device=dev),
                torch.zeros(1, dtype=torch.int32, device=dev))
        return self._update_buffers

    def _batch_count(self):
        seg = self.replay
        if seg.batch_size is None:
            return seg.batch_iterations
        total = seg.max_size * seg.num_workers * distributed.world()
        return seg.batch_iterations * -(-total // seg.batch_size)

    def _enqueue_update(self, stats, stop):
        """All kernels of one update, no host synchronisation (CUDA-graph capturable)."""
        stats.zero_()
        self._evaluate()
        flat = self.replay.get_full('observations', 'actions', 'advantages', 'log_probs',
                                    'returns')
        total = flat['observations'].shape[0]
        # one policy-gradient step on the full batch (a2c.py:107-114)
        self.actor_updater.launch(flat['observations'], flat['actions'], flat['advantages'],
                                  flat['log_probs'], None, total, stats[0, 0],
                                  rows_global=total * distributed.world())
        # several value-regression steps (a2c.py:116-121)
        for j, (idx, rows, rows_global) in enumerate(self.replay.index_batches()):
            self.critic_updater.launch(flat['observations'], flat['returns'], idx, rows,
                                       stats[j + 1, 1], rows_global=rows_global)
        if self.model.observation_normalizer:
            self.model.observation_normalizer.update()

    def _report(self, host):
        for k, v in self.actor_updater.infos(host[0, 0]).items():
            logger.store('actor/' + k, v)
        for j in range(1, host.shape[0]):
            for k, v in self.critic_updater.infos(host[j, 1]).items():
                logger.store('critic/' + k, v)

    extra_stat_rows = 1      # A2C: row 0 holds the single policy-gradient step

    def _update(self):
        stats, stop = self._buffers_for_update(self._batch_count() + self.extra_stat_rows)
        if self._graphable():
            if self._update_graph is None:
                self._update_graph = graphs.CapturedSection(
                    lambda: self._enqueue_update(stats, stop))
            self._update_graph()
            self.replay.index = 0
        else:
            self._enqueue_update(stats, stop)
        self._report(kernels.to_host(stats))
