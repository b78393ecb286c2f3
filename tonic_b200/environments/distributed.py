"""Device-resident vector environment.

Replaces the reference's worker grid -- `Sequential` (a Python loop over M
environments, tonic/environments/distributed.py:8-66) and `Parallel` (P forked
processes exchanging pickled arrays, :69-155) -- by one fused sm_100a kernel
that steps all N = P*M environments resident in HBM (csrc/env_step.cu).  The
object keeps the reference's protocol: `initialize(seed)`, `start()`,
`step(actions) -> (observations, infos)`, attributes `observation_space`,
`action_space`, `name`, `max_episode_steps`.

Worker i is seeded `seed + i` exactly like distributed.py:18-20,109; with several
ranks (one process per GPU) rank r owns workers [r*N/W, (r+1)*N/W), the same
contiguous split as `np.split` at distributed.py:137.

Arrays handed back are views of device buffers that stay valid until the next
`step` (the reference returns fresh host arrays).  If `step` is given a numpy
array the results are copied to the host and returned as numpy arrays with the
reference's dtypes (float32 / bool), which is the drop-in path for host code.
"""

import ctypes

import numpy as np
import torch

from .. import _lib, kernels
from .._lib import ptr


class DeviceVectorEnvironment:
    def __init__(self, spec, workers, first_worker=0, episode_log_capacity=1 << 20):
        self.spec = spec
        self.workers = int(workers)
        self.first_worker = int(first_worker)
        self.observation_space = spec.observation_space
        self.action_space = spec.action_space
        self.name = spec.name
        self.max_episode_steps = spec.max_episode_steps
        self.episode_log_capacity = int(episode_log_capacity)
        self.started = False

    def __len__(self):
        return self.workers

    def initialize(self, seed):
        dev = kernels.device()
        N, O = self.workers, self.spec.observation_size
        cap = self.episode_log_capacity
        self.seed = int(seed)
        self.state = torch.zeros(N, O, dtype=torch.float32, device=dev)
        self.lengths = torch.zeros(N, dtype=torch.int32, device=dev)
        self.episodes = torch.zeros(N, dtype=torch.int32, device=dev)
        self.scores = torch.zeros(N, dtype=torch.float64, device=dev)
        self.episode_scores = torch.zeros(cap, dtype=torch.float64, device=dev)
        self.episode_lengths = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.episode_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._episodes_read = 0
        # output buffers (views handed to the caller)
        tf = bool(getattr(self.spec, 'time_feature', False))
        self.time_feature = tf
        task = int(getattr(self.spec, 'task_id', 0))
        self.state64 = torch.zeros(N, 2, dtype=torch.float64, device=dev) if task else None
        # one packed block (regions 256-byte aligned): the host protocol fetches all five results
        # with a single device->host copy
        sizes = [N * (O + tf), N * (O + tf), N, N, N]
        offsets, total = [], 0
        for size in sizes:
            offsets.append(total)
            total = (total + size + 63) // 64 * 64
        self._out_block = torch.zeros(total, dtype=torch.float32, device=dev)
        self._out_regions = list(zip(offsets, sizes, [(N, O + tf), (N, O + tf), (N,), (N,), (N,)]))
        (self.observations, self.next_observations, self.rewards, self.resets,
         self.terminations) = (self._out_block[o:o + n].view(shape) for o, n, shape in self._out_regions)
        self.struct = _lib.TbEnv(
            n_envs=N, obs_dim=O, act_dim=self.spec.action_size,
            max_episode_steps=self.max_episode_steps, seed=self.seed,
            first_worker=self.first_worker, d_state=ptr(self.state),
            d_length=ptr(self.lengths), d_episode=ptr(self.episodes), d_score=ptr(self.scores),
            d_ep_scores=ptr(self.episode_scores), d_ep_lengths=ptr(self.episode_lengths),
            d_ep_count=ptr(self.episode_count), log_cap=cap, time_feature=int(tf),
            time_low=-1.0, time_high=1.0, task=task, d_state64=ptr(self.state64))

    def start(self, host=False):
        """Resets every environment; returns the first observations [N, O]."""
        assert not self.started or True
        self.started = True
        _lib.call('tb_env_start', ctypes.byref(self.struct), ptr(self.observations),
                  kernels.stream())
        if host:
            return kernels.to_host(self.observations)
        return self.observations

    def step_into(self, actions, observations, next_observations, rewards, resets, terminations):
        """Device-only step writing into caller-provided buffers (segment rows)."""
        _lib.call('tb_env_step', ctypes.byref(self.struct), ptr(actions), ptr(observations),
                  ptr(next_observations), ptr(rewards), ptr(resets), ptr(terminations),
                  kernels.stream())

    HOST_SLOTS = 3      # arrays returned by step() stay valid for this many further calls

    def _step_host(self, actions):
        """environment.step(numpy) -> numpy: pinned copy-in, then H2D + the step kernel as ONE
        captured graph, ONE device->host copy of the packed result block into a rotating pinned
        slot, one synchronisation.  The returned arrays are views of that slot (no host copy): they
        are overwritten HOST_SLOTS calls later -- the reference's agents copy what they keep
        (torch/agents/a2c.py:48), and so does everything in this package."""
        from .. import graphs
        if getattr(self, '_bridge', None) is None:
            self._bridge, self._host_section = kernels.HostBridge(), None
        b = self._bridge
        pin_act, dev_act = b.load('actions', actions)
        if self._host_section is None:
            def body():
                dev_act.copy_(pin_act, non_blocking=True)
                self.step_into(dev_act, self.observations, self.next_observations, self.rewards,
                               self.resets, self.terminations)
            self._host_section = graphs.CapturedSection(body)
            self._host_slots = [torch.empty(self._out_block.numel(), dtype=torch.float32).pin_memory()
                                for _ in range(self.HOST_SLOTS)]
            self._host_views = [[slot.numpy()[o:o + n].reshape(shape) for o, n, shape in self._out_regions]
                                for slot in self._host_slots]
            self._host_slot = 0
            self._host_bytes = 4 * sum(n for _, n, _ in self._out_regions)
        self._host_section()
        self._host_slot = (self._host_slot + 1) % self.HOST_SLOTS
        self._host_slots[self._host_slot].copy_(self._out_block, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        kernels.transfers['d2h'] += self._host_bytes
        obs, next_obs, rewards, resets, terminations = self._host_views[self._host_slot]
        return obs, dict(observations=next_obs, rewards=rewards, resets=resets.astype(np.bool_),
                         terminations=terminations.astype(np.bool_))

    def step(self, actions):
        host = not isinstance(actions, torch.Tensor) or not actions.is_cuda
        from .. import config
        if host and config.graphs and config.noise == 'device':
            actions = np.asarray(actions, np.float32)
            assert actions.shape == (self.workers, self.spec.action_size), actions.shape
            return self._step_host(actions)
        actions = kernels.to_device(actions)
        assert actions.shape == (self.workers, self.spec.action_size), actions.shape
        self.step_into(actions, self.observations, self.next_observations, self.rewards,
                       self.resets, self.terminations)
        if host:
            # one synchronisation for the five result arrays (pinned staging)
            obs, next_obs, rewards, resets, terminations = kernels.to_host(
                self.observations, self.next_observations, self.rewards, self.resets,
                self.terminations)
            infos = dict(observations=next_obs, rewards=rewards, resets=resets.astype(np.bool_),
                         terminations=terminations.astype(np.bool_))
            return obs, infos
        infos = dict(observations=self.next_observations, rewards=self.rewards,
                     resets=self.resets, terminations=self.terminations)
        return self.observations, infos

    def finished_episodes(self):
        """(scores, lengths) of the episodes finished since the last call
        (what trainer.py:64-71 collects with a per-worker Python loop)."""
        total = int(self.episode_count.item())
        new = total - self._episodes_read
        cap = self.episode_log_capacity
        if new <= 0:
            return np.zeros(0), np.zeros(0, int)
        new = min(new, cap)
        idx = (torch.arange(total - new, total, device=self.state.device) % cap)
        scores = kernels.to_host(self.episode_scores[idx])
        lengths = kernels.to_host(self.episode_lengths[idx]).astype(int)
        self._episodes_read = total
        return scores, lengths

    def render(self, *args, **kwargs):
        raise NotImplementedError('synthetic device environments have no renderer')


class ClassicSpec:
    """Device description of a closed-form classic-control task built by
    `environments.Gym(name)` (environments/classic.py): same spaces / name / time limit, the
    dynamics run in csrc/classic_env.cu (bit-identical to the numpy classes)."""

    def __init__(self, wrapped):
        self.task_id = int(wrapped.task_id)
        self.name = wrapped.name
        self.max_episode_steps = int(wrapped.max_episode_steps)
        self.time_feature = bool(wrapped.time_feature)
        self.observation_space = wrapped.observation_space
        self.action_space = wrapped.action_space
        self.observation_size = wrapped.environment.observation_space.shape[0]
        self.action_size = wrapped.action_space.shape[0]


def _classic_on_device():
    """TONIC_B200_CLASSIC=host keeps Gym(name) tasks on the host worker grid."""
    import os
    return os.environ.get('TONIC_B200_CLASSIC', 'device') != 'host' and torch.cuda.is_available()


def distribute(environment_builder, worker_groups=1, workers_per_group=1, single=False):
    """Same signature as the reference's `distribute`
    (tonic/environments/distributed.py:158-172).  `worker_groups *
    workers_per_group` environments are created in total; under torchrun they are
    sharded contiguously over the ranks (one GPU each).  `single=True` builds ONE
    environment per process whatever the number of ranks (the test environment of
    train.py:88-91)."""
    spec = environment_builder()
    if getattr(spec, 'task_id', 0) and _classic_on_device():
        # Gym(name) for a closed-form task: stepped by the device kernel (csrc/classic_env.cu)
        spec = ClassicSpec(spec)
    if not hasattr(spec, 'observation_size'):
        # a host environment (Gym / dm_control / user code): the reference's worker grid on the
        # host, feeding the device learner through numpy arrays (tonic_b200/environments/host.py)
        from . import host
        if single:
            worker_groups = workers_per_group = 1
        return host.distribute_host(environment_builder, worker_groups, workers_per_group)
    if single:
        return DeviceVectorEnvironment(spec, 1)
    total = int(worker_groups) * int(workers_per_group)
    rank, world = 0, 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    if total % world:
        raise ValueError(f'{total} workers do not split over {world} ranks')
    local = total // world
    return DeviceVectorEnvironment(spec, local, first_worker=rank * local)
