"""Closed-form classic-control tasks as host (numpy) environments.

`gym` is not available where this backend is built, so the tasks the reference's README uses
for its CPU quick-start (`tonic.environments.Gym('Pendulum-v1')`, BASELINE.json configs[0])
are restated here from Gym's published dynamics -- parity with Gym itself is NOT pinned by any
fixture -- and wrapped the way `build_environment` wraps a Gym task
(reference environments/builders.py:43-78): the TimeLimit is removed and remembered as
`max_episode_steps` (time-outs then reset without terminating, distributed.py:39-40), actions
are rescaled from [-1, 1]^n to the task's bounds (ActionRescaler, wrappers.py:7-22) and an
optional time feature is appended (wrappers.py:25-54).  They run on the host worker grid
(`environments/host.py`) and feed the device learner through numpy arrays.
"""

import numpy as np

from .builders import Space


class Pendulum:
    """Gym `Pendulum-v0/v1`: swing a torque-limited pendulum upright.
    state (theta, theta_dot); observation (cos theta, sin theta, theta_dot);
    theta_dot' = clip(theta_dot + (3 g / (2 l) sin theta + 3 / (m l^2) u) dt, +-8);
    theta' = theta + theta_dot' dt; reward = -(wrap(theta)^2 + 0.1 theta_dot^2 + 0.001 u^2),
    never terminates; 200-step time limit."""

    max_speed, max_torque, dt, g, m, length = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0
    default_max_episode_steps = 200

    def __init__(self):
        high = np.array([1.0, 1.0, self.max_speed], np.float32)
        self.observation_space = Space(3)
        self.observation_space.low, self.observation_space.high = -high, high
        self.action_space = Space(1, -self.max_torque, self.max_torque)
        self.np_random = np.random.RandomState()
        self.state = np.zeros(2)

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)

    def reset(self):
        self.state = self.np_random.uniform(low=[-np.pi, -1.0], high=[np.pi, 1.0])
        return self._observation()

    def step(self, action):
        theta, theta_dot = self.state
        u = float(np.clip(action, -self.max_torque, self.max_torque)[0])
        wrapped = ((theta + np.pi) % (2 * np.pi)) - np.pi
        cost = wrapped ** 2 + 0.1 * theta_dot ** 2 + 0.001 * u ** 2
        theta_dot = theta_dot + (3 * self.g / (2 * self.length) * np.sin(theta)
                                 + 3.0 / (self.m * self.length ** 2) * u) * self.dt
        theta_dot = float(np.clip(theta_dot, -self.max_speed, self.max_speed))
        theta = theta + theta_dot * self.dt
        self.state = np.array([theta, theta_dot])
        return self._observation(), -cost, False, {}

    def _observation(self):
        theta, theta_dot = self.state
        return np.array([np.cos(theta), np.sin(theta), theta_dot], np.float32)

    def render(self, *args, **kwargs):
        return None


class _Wrapped:
    """What `build_environment` returns: rescaled actions, optional time feature, the time
    limit as an attribute (builders.py:56-76)."""

    def __init__(self, environment, name, max_episode_steps, time_feature):
        self.environment, self.name = environment, name
        self.max_episode_steps = max_episode_steps
        self.time_feature = time_feature
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


TASKS = {'Pendulum-v0': Pendulum, 'Pendulum-v1': Pendulum}


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
