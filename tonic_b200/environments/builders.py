"""Environment builders.

The reference builds Gym / PyBullet / dm_control environments on the host
(tonic/environments/builders.py:12-78); those physics engines are outside the
hot path.  The device backend steps synthetic
continuous-control environments of the same SHAPE (observation / action sizes,
time limit, unit action box after ActionRescaler) entirely on the GPU, so a
builder here only describes the environment; the dynamics live in
csrc/env_step.cu.
"""

import numpy as np


class Space:
    """Box-like space: what agents read from an environment (shape/low/high)."""

    def __init__(self, size, low=-1.0, high=1.0):
        self.shape = (int(size),)
        self.low = np.full(size, low, np.float32)
        self.high = np.full(size, high, np.float32)
        self.dtype = np.dtype(np.float32)


class SynthControl:
    """Description of the synthetic SynthControl(O, A) task.

    Presets mirror BASELINE.json's shapes: 'HalfCheetah' (17, 6), 'Humanoid'
    (376, 17), 'Ant' (111, 8).  `max_episode_steps` plays the role of the
    TimeLimit the reference reads at builders.py:56-59 (time-outs reset the
    episode without terminating it, distributed.py:39-40).
    """

    PRESETS = {'HalfCheetah': (17, 6), 'Humanoid': (376, 17), 'Ant': (111, 8),
               'Pendulum': (3, 1)}

    def __init__(self, name='HalfCheetah', observation_size=None, action_size=None,
                 max_episode_steps=1000, time_feature=False):
        if observation_size is None or action_size is None:
            observation_size, action_size = self.PRESETS[name]
        self.name = f'SynthControl-{name}' if name in self.PRESETS else str(name)
        self.observation_size = int(observation_size)
        self.action_size = int(action_size)
        # time_feature: what build_environment(time_feature=True) adds (reference
        # environments/builders.py:65-68, wrappers.py:25-54): one more observation column
        self.time_feature = bool(time_feature)
        self.observation_space = Space(observation_size + int(self.time_feature), -np.inf, np.inf)
        self.action_space = Space(action_size)          # wrappers.py:14-15: [-1, 1]^n
        self.max_episode_steps = int(max_episode_steps)
