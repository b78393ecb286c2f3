"""Host-side vector environments: the bridge for environments that cannot live on the GPU.

The device backend steps its synthetic tasks inside CUDA kernels
(`DeviceVectorEnvironment`), but the learner only needs the reference's vector-environment
protocol: `initialize(seed)`, `start() -> observations [N, O]`, `step(actions [N, A]) ->
(observations, dict(observations, rewards, resets, terminations))`.  The classes below give
that protocol to ANY gym-like host environment (Gym, dm_control wrappers, user code), so that
`tonic_b200` agents can be trained on them through the host-array path of
`agent.step / agent.update` (pinned staging, see tonic_b200/kernels.py).

Semantics restated from the reference (tonic/environments/distributed.py):
  * worker j is seeded `seed + j`                                     (:18-20, :109)
  * `lengths += 1`; `reset = termination or lengths == max_episode_steps`; a time-out resets
    the episode WITHOUT being a termination                            (:36-40)
  * after a reset the returned acting observation is the new episode's first one, while
    `infos['observations']` keeps the transition's last observation   (:41-50)
  * `Parallel`: `worker_groups` forked daemon processes, each stepping `workers_per_group`
    environments; actions are split contiguously (`np.split`), results concatenated in group
    order; every step is a synchronous scatter / gather                (:81-155)
dtypes: observations / rewards float32, resets / terminations bool   (:52-58).
"""

import multiprocessing

import numpy as np


class HostSequential:
    """`workers` environments stepped one after the other in this process."""

    def __init__(self, environment_builder, max_episode_steps, workers, first_worker=0):
        self.environments = [environment_builder() for _ in range(workers)]
        self.max_episode_steps = max_episode_steps
        self.first_worker = first_worker
        head = self.environments[0]
        self.observation_space = head.observation_space
        self.action_space = head.action_space
        self.name = getattr(head, 'name', type(head).__name__)

    def __len__(self):
        return len(self.environments)

    def initialize(self, seed):
        for j, environment in enumerate(self.environments):
            environment.seed(seed + self.first_worker + j)

    def start(self):
        self.lengths = np.zeros(len(self.environments), int)
        return np.array([environment.reset() for environment in self.environments], np.float32)

    def step(self, actions):
        acting, last, rewards, resets, terminations = [], [], [], [], []
        for j, environment in enumerate(self.environments):
            observation, reward, termination, _ = environment.step(actions[j])
            self.lengths[j] += 1
            reset = termination or self.lengths[j] == self.max_episode_steps
            last.append(observation)
            rewards.append(reward)
            resets.append(reset)
            terminations.append(termination)
            if reset:
                observation = environment.reset()
                self.lengths[j] = 0
            acting.append(observation)
        infos = dict(observations=np.array(last, np.float32), rewards=np.array(rewards, np.float32),
                     resets=np.array(resets, np.bool_), terminations=np.array(terminations, np.bool_))
        return np.array(acting, np.float32), infos

    def render(self, mode='human', *args, **kwargs):
        frames = [environment.render(mode=mode, *args, **kwargs) for environment in self.environments]
        if mode != 'human':
            return np.array(frames)

    def close(self):
        pass


def _group_process(builder, max_episode_steps, workers, first_worker, seed, pipe):
    """Body of one worker group: a HostSequential driven through a pipe."""
    group = HostSequential(builder, max_episode_steps, workers, first_worker)
    group.initialize(seed)
    pipe.send(('ready', None))
    while True:
        command, payload = pipe.recv()
        if command == 'start':
            pipe.send(group.start())
        elif command == 'step':
            pipe.send(group.step(payload))
        elif command == 'close':
            pipe.close()
            return


class HostParallel:
    """`worker_groups` forked processes x `workers_per_group` environments each."""

    def __init__(self, environment_builder, worker_groups, workers_per_group, max_episode_steps):
        self.builder = environment_builder
        self.worker_groups, self.workers_per_group = worker_groups, workers_per_group
        self.max_episode_steps = max_episode_steps
        probe = environment_builder()
        self.observation_space, self.action_space = probe.observation_space, probe.action_space
        self.name = getattr(probe, 'name', type(probe).__name__)
        self.processes, self.pipes = [], []

    def __len__(self):
        return self.worker_groups * self.workers_per_group

    def initialize(self, seed):
        context = multiprocessing.get_context('fork')
        for g in range(self.worker_groups):
            ours, theirs = context.Pipe()
            process = context.Process(
                target=_group_process, daemon=True,
                args=(self.builder, self.max_episode_steps, self.workers_per_group,
                      g * self.workers_per_group, seed, theirs))
            process.start()
            self.processes.append(process)
            self.pipes.append(ours)
        for pipe in self.pipes:
            assert pipe.recv()[0] == 'ready'

    def start(self):
        for pipe in self.pipes:
            pipe.send(('start', None))
        return np.concatenate([pipe.recv() for pipe in self.pipes])

    def step(self, actions):
        for pipe, block in zip(self.pipes, np.split(np.asarray(actions), self.worker_groups)):
            pipe.send(('step', block))
        results = [pipe.recv() for pipe in self.pipes]
        observations = np.concatenate([r[0] for r in results])
        infos = {key: np.concatenate([r[1][key] for r in results]) for key in results[0][1]}
        return observations, infos

    def close(self):
        for pipe in self.pipes:
            pipe.send(('close', None))
        for process in self.processes:
            process.join(timeout=5)
        self.processes, self.pipes = [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def distribute_host(environment_builder, worker_groups=1, workers_per_group=1):
    """Reference `distribute` (distributed.py:158-172) for host environments: a dummy
    environment provides `max_episode_steps`; fewer than two groups run in-process."""
    probe = environment_builder()
    max_episode_steps = probe.max_episode_steps
    del probe
    if worker_groups < 2:
        return HostSequential(environment_builder, max_episode_steps, workers_per_group)
    return HostParallel(environment_builder, worker_groups, workers_per_group, max_episode_steps)
