"""The agent protocol (boundary contract, tonic/agents/agent.py:4-34): the seven
methods the trainer calls.  Kept verbatim in meaning; arrays may be numpy
(host, drop-in mode) or CUDA tensors (device mode)."""

import abc


class Agent(abc.ABC):
    def initialize(self, observation_space, action_space, seed=None):
        pass

    @abc.abstractmethod
    def step(self, observations, steps):
        """Actions for the training environments."""

    def update(self, observations, rewards, resets, terminations, steps):
        """Receives the transitions that followed the last `step`."""

    @abc.abstractmethod
    def test_step(self, observations, steps):
        """Actions for the test environment."""

    def test_update(self, observations, rewards, resets, terminations, steps):
        pass

    def save(self, path):
        pass

    def load(self, path):
        pass
