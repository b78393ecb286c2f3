"""Closed-form classic-control tasks as host (numpy) environments.

`gym` is not available where this backend is built, so the tasks the reference's README uses
for its CPU quick-start (`tonic.environments.Gym('Pendulum-v1')`, BASELINE.json configs[0])
are restated here from Gym's published dynamics -- parity with Gym itself is NOT pinned by any
fixture; in particular the reset states come from a counter-based hash
stream (same distributions as Gym's `np_random.uniform`, different numbers) and sin / cos from
`portable_math` so that the device kernels (csrc/classic_env.cu) reproduce these classes bit for
bit -- and wrapped the way `build_environment` wraps a Gym task
(reference environments/builders.py:43-78): the TimeLimit is removed and remembered as
`max_episode_steps` (time-outs then reset without terminating, distributed.py:39-40), actions
are rescaled from [-1, 1]^n to the task's bounds (ActionRescaler, wrappers.py:7-22) and an
optional time feature is appended (wrappers.py:25-54).  They run on the host worker grid
(`environments/host.py`) and feed the device learner through numpy arrays.
"""

import numpy as np

from . import portable_math
from .builders import Space


class _Task:
    """Common part of the restated tasks: float64 state, counter-based resets (24 hashed bits
    of (seed, episode, coordinate), `portable_math.reset_uniform`: the stream the device kernel
    reproduces), arithmetic restricted to individually rounded float64 operations."""

    task_id = 0
    state_size = 2

    def __init__(self):
        self._seed, self._episode = 0, 0
        self.state = np.zeros(self.state_size)

    def seed(self, seed=None):
        self._seed, self._episode = int(seed or 0), 0

    def _uniform(self, coordinate):
        return float(portable_math.reset_uniform(self._seed, self._episode, coordinate))

    def render(self, *args, **kwargs):
        return None


class Pendulum(_Task):
    """Gym `Pendulum-v0/v1`: swing a torque-limited pendulum upright.
    state (theta, theta_dot); observation (cos theta, sin theta, theta_dot);
    theta_dot' = clip(theta_dot + (3 g / (2 l) sin theta + 3 / (m l^2) u) dt, +-8);
    theta' = theta + theta_dot' dt; reward = -(wrap(theta)^2 + 0.1 theta_dot^2 + 0.001 u^2),
    never terminates; 200-step time limit.  Resets: theta ~ U(-pi, pi), theta_dot ~ U(-1, 1)."""

    max_speed, max_torque, dt, g, m, length = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0
    default_max_episode_steps = 200
    task_id = 1

    def __init__(self):
        super().__init__()
        high = np.array([1.0, 1.0, self.max_speed], np.float32)
        self.observation_space = Space(3)
        self.observation_space.low, self.observation_space.high = -high, high
        self.action_space = Space(1, -self.max_torque, self.max_torque)

    def reset(self):
        self.state = np.array([-np.pi + (2 * np.pi) * self._uniform(0), -1.0 + 2.0 * self._uniform(1)])
        self._episode += 1
        return self._observation()

    def step(self, action):
        theta, theta_dot = (float(v) for v in self.state)
        u = float(np.clip(np.asarray(action, np.float64).ravel()[0], -self.max_torque, self.max_torque))
        wrapped = ((theta + np.pi) % (2 * np.pi)) - np.pi
        cost = (wrapped * wrapped + 0.1 * (theta_dot * theta_dot)) + 0.001 * (u * u)
        sin_theta = float(portable_math.sincos(theta)[0])
        theta_dot = theta_dot + (15.0 * sin_theta + 3.0 * u) * self.dt     # 3 g / (2 l) = 15, 3 / (m l^2) = 3
        theta_dot = min(max(theta_dot, -self.max_speed), self.max_speed)
        theta = theta + theta_dot * self.dt
        self.state = np.array([theta, theta_dot])
        return self._observation(), -cost, False, {}

    def _observation(self):
        theta, theta_dot = self.state
        sin_theta, cos_theta = portable_math.sincos(theta)
        return np.array([cos_theta, sin_theta, theta_dot], np.float32)


class MountainCarContinuous(_Task):
    """Gym `MountainCarContinuous-v0`: state (position, velocity) = observation;
    velocity' = clip(velocity + 0.0015 force - 0.0025 cos(3 position), +-0.07);
    position' = clip(position + velocity', -1.2, 0.6) (velocity' = 0 at the left wall);
    terminates at position' >= 0.45 with reward 100; reward -= 0.1 force^2; 999-step limit.
    Resets: position ~ U(-0.6, -0.4), velocity 0."""

    default_max_episode_steps = 999
    task_id = 2

    def __init__(self):
        super().__init__()
        self.observation_space = Space(2)
        self.observation_space.low = np.array([-1.2, -0.07], np.float32)
        self.observation_space.high = np.array([0.6, 0.07], np.float32)
        self.action_space = Space(1, -1.0, 1.0)

    def reset(self):
        self.state = np.array([-0.6 + 0.2 * self._uniform(0), 0.0])
        self._episode += 1
        return self.state.astype(np.float32)

    def step(self, action):
        position, velocity = (float(v) for v in self.state)
        force = float(np.clip(np.asarray(action, np.float64).ravel()[0], -1.0, 1.0))
        cos3 = float(portable_math.sincos(3.0 * position)[1])
        velocity = velocity + (force * 0.0015 - 0.0025 * cos3)
        velocity = min(max(velocity, -0.07), 0.07)
        position = position + velocity
        position = min(max(position, -1.2), 0.6)
        if position == -1.2 and velocity < 0:
            velocity = 0.0
        done = bool(position >= 0.45 and velocity >= 0.0)
        reward = (100.0 if done else 0.0) - (force * force) * 0.1
        self.state = np.array([position, velocity])
        return self.state.astype(np.float32), reward, done, {}


class _Wrapped:
    """What `build_environment` returns: rescaled actions, optional time feature, the time
    limit as an attribute (builders.py:56-76)."""

    def __init__(self, environment, name, max_episode_steps, time_feature):
        self.environment, self.name = environment, name
        self.max_episode_steps = max_episode_steps
        self.time_feature = time_feature
        # what the device backend needs to run this task as a kernel (csrc/classic_env.cu)
        self.task_id = environment.task_id
        low, high = environment.action_space.low, environment.action_space.high
        self.scale, self.bias = (high - low) / 2, (high + low) / 2          # wrappers.py:10-13
        self.action_space = Space(len(low))                                 # [-1, 1]^n
        self.observation_space = environment.observation_space
        if time_feature:
            size = environment.observation_space.shape[0] + 1
            self.observation_space = Space(size)
            self.observation_space.low = np.append(environment.observation_space.low, -1).astype(np.float32)
            self.observation_space.high = np.append(environment.observation_space.high, 1).astype(np.float32)
        self.steps = 0

    def seed(self, seed=None):
        self.environment.seed(seed)

    def _timed(self, observation):
        if not self.time_feature:
            return observation
        value = -1 + 2 * (self.steps / self.max_episode_steps) if self.steps else -1
        return np.append(observation, value)

    def reset(self):
        self.steps = 0
        return self._timed(self.environment.reset())

    def step(self, action):
        action = self.bias + self.scale * np.clip(action, -1, 1)            # wrappers.py:21-22
        observation, reward, termination, info = self.environment.step(action)
        self.steps += 1
        return self._timed(observation), reward, termination, info

    def render(self, *args, **kwargs):
        return self.environment.render(*args, **kwargs)


TASKS = {'Pendulum-v0': Pendulum, 'Pendulum-v1': Pendulum,
         'MountainCarContinuous-v0': MountainCarContinuous}


def Gym(name, terminal_timeouts=False, time_feature=False, max_episode_steps='default',
        scaled_actions=True):
    """`tonic.environments.Gym(name, ...)` (reference builders.py:12-16,43-78) for the tasks
    restated in this module."""
    if name not in TASKS:
        raise NotImplementedError(
            f'{name}: only {sorted(TASKS)} are restated here; other Gym tasks need the gym package '
            '(wrap them yourself and pass the builder to tonic_b200.environments.distribute)')
    if terminal_timeouts:
        raise NotImplementedError('terminal_timeouts=True (time-outs as terminations) is not restated')
    if not scaled_actions:
        raise NotImplementedError('scaled_actions=False is not restated')
    task = TASKS[name]()
    if max_episode_steps == 'default':
        max_episode_steps = task.default_max_episode_steps
    return _Wrapped(task, name, int(max_episode_steps), bool(time_feature))
