"""Proximal Policy Optimization on the device (reference:
tonic/torch/agents/ppo.py:7-67).

The reference decides on the host, after every minibatch, whether the actor
keeps training (`kl > 0.015`, ppo.py:45-46).  Here the decision lives in a device
flag written by the Adam kernel and read by the actor kernels of the following
minibatches, so the E x (T*N/B) minibatch updates are enqueued back to back
and the host reads all statistics once at the end.
"""

import torch

from ... import _lib, kernels
from ...utils import logger
from .. import updaters
from . import a2c


class PPO(a2c.A2C):
    def __init__(self, model=None, replay=None, actor_updater=None, critic_updater=None):
        actor_updater = actor_updater or updaters.ClippedRatio()
        super().__init__(model=model, replay=replay, actor_updater=actor_updater,
                         critic_updater=critic_updater)

    extra_stat_rows = 0

    def _enqueue_update(self, stats, stop):
        """All kernels of one update, no host synchronisation (CUDA-graph capturable):
        the KL decision of ppo.py:45-46 is taken by the Adam kernel on the device."""
        stats.zero_()
        stop.zero_()
        self._evaluate()
        flat = self.replay.get_full('observations', 'actions', 'advantages', 'log_probs',
                                    'returns')
        for j, (idx, rows, rows_global) in enumerate(self.replay.index_batches()):
            self.actor_updater.launch(flat['observations'], flat['actions'],
                                      flat['advantages'], flat['log_probs'], idx, rows,
                                      stats[j, 0], stop=stop, rows_global=rows_global)
            self.critic_updater.launch(flat['observations'], flat['returns'], idx, rows,
                                       stats[j, 1], rows_global=rows_global)
        if self.model.observation_normalizer:
            self.model.observation_normalizer.update()

    def _report(self, host):
        actor_iterations = 0
        for j in range(host.shape[0]):
            if host[j, 0, _lib.STAT_ROWS] > 0:      # the actor was still training
                actor_iterations += 1
                for k, v in self.actor_updater.infos(host[j, 0]).items():
                    logger.store('actor/' + k, v)
            for k, v in self.critic_updater.infos(host[j, 1]).items():
                logger.store('critic/' + k, v)
        logger.store('actor/iterations', actor_iterations)
        logger.store('critic/iterations', host.shape[0])
