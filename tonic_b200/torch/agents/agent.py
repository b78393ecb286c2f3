"""Base class of the device agents (reference: tonic/torch/agents/agent.py:10-26):
seeding of numpy / random / torch, and `.pt` checkpoints holding
`model.state_dict()` with the reference's key layout."""

import os
import random

import numpy as np
import torch

from ... import agents
from ...utils import logger


class Agent(agents.Agent):
    def initialize(self, seed=None):
        self.seed = seed
        if seed is not None:
            np.random.seed(seed)
            random.seed(seed)
            torch.manual_seed(seed)

    def save(self, path):
        path = path + '.pt'
        logger.log(f'\nSaving weights to {path}')
        os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
        torch.save(self.model.state_dict(), path)

    def load(self, path):
        path = path + '.pt'
        logger.log(f'\nLoading weights from {path}')
        self.model.load_state_dict(torch.load(path))
