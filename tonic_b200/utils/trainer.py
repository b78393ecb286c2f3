"""Training loop with the reference's constructor, `initialize`, `run` and log
keys (tonic/utils/trainer.py:9-146).

Two drivers behind `run()`:

* stepwise -- the reference's call sequence, one `agent.step` /
  `environment.step` / `agent.update` per vector step; works with host (numpy)
  or device arrays and is the drop-in path;
* fused -- when the environment is device resident and the agent offers
  `rollout`, whole stretches of vector steps are enqueued without returning to
  the host; episode scores / lengths come from the environment's device-side
  episode log and action statistics from a reduction kernel instead of the
  per-worker Python loop at trainer.py:64-71.  Step counts, epoch boundaries,
  test / checkpoint cadence and logged keys are unchanged.
"""

import os
import time

import numpy as np
import torch

from .. import kernels
from . import logger


class Trainer:
    def __init__(self, steps=int(1e7), epoch_steps=int(2e4), save_steps=int(5e5),
                 test_episodes=5, show_progress=True, replace_checkpoint=False):
        self.max_steps = steps
        self.epoch_steps = epoch_steps
        self.save_steps = save_steps
        self.test_episodes = test_episodes
        self.show_progress = show_progress
        self.replace_checkpoint = replace_checkpoint

    def initialize(self, agent, environment, test_environment=None):
        self.agent = agent
        self.environment = environment
        self.test_environment = test_environment

    # ------------------------------------------------------------------ run
    def run(self):
        fused = (hasattr(self.agent, 'rollout') and hasattr(self.environment, 'step_into')
                 and getattr(self.agent, 'can_rollout', lambda env: True)(self.environment))
        self.start_time = self.last_epoch_time = time.time()
        self.steps = self.epoch_step_count = self.epochs = self.episodes = 0
        self.steps_since_save = 0
        (self._run_fused if fused else self._run_stepwise)()

    def _run_fused(self):
        env, agent = self.environment, self.agent
        env.start()
        workers = env.workers
        stats = kernels.ArrayStats()
        while True:
            # vector steps until the next epoch end, checkpoint or the end of training,
            # whichever comes first (trainer.py:73-112 checks all three after every step)
            to_go = min(self.epoch_steps - self.epoch_step_count,
                        self.max_steps - self.steps,
                        self.save_steps - self.steps_since_save)
            budget = max(1, -(-to_go // workers))
            done = agent.rollout(env, budget, steps=self.steps, action_stats=stats)
            self._advance(done * workers)
            if self.show_progress:
                logger.show_progress(self.steps, self.epoch_steps, self.max_steps)
            if self.epoch_step_count >= self.epoch_steps:
                count, total, total_sq, low, high = stats.read()
                assert not np.isnan(total), 'NaN in the actions'      # trainer.py:45
                logger.store_aggregate('train/action', count, total, total_sq, low, high,
                                       stats=True, items=stats.items)
                stats.reset()
                scores, lengths = env.finished_episodes()
                self.episodes += len(scores)
                if len(scores):       # one value per finished episode (trainer.py:66-68)
                    logger.store('train/episode_score', scores, stats=True, items=len(scores))
                    logger.store('train/episode_length', lengths, stats=True, items=len(lengths))
                self._end_epoch(workers)
            if self._checkpoint_and_stop():
                break

    def _run_stepwise(self):
        env, agent = self.environment, self.agent
        observations = env.start()
        workers = len(observations)
        scores = np.zeros(workers)
        lengths = np.zeros(workers, int)
        while True:
            actions = agent.step(observations, self.steps)
            host_actions = actions.detach().cpu().numpy() if torch.is_tensor(actions) else actions
            assert not np.isnan(host_actions.sum())
            logger.store('train/action', host_actions, stats=True)
            observations, infos = env.step(actions)
            agent.update(**infos, steps=self.steps)
            rewards, resets = infos['rewards'], infos['resets']
            if torch.is_tensor(rewards):
                rewards, resets = rewards.cpu().numpy(), resets.cpu().numpy() != 0
            scores += rewards
            lengths += 1
            self._advance(workers)
            if self.show_progress:
                logger.show_progress(self.steps, self.epoch_steps, self.max_steps)
            finished = np.flatnonzero(resets)
            if len(finished):
                logger.store('train/episode_score', scores[finished], stats=True,
                             items=len(finished))
                logger.store('train/episode_length', lengths[finished], stats=True,
                             items=len(finished))
                scores[finished] = 0
                lengths[finished] = 0
                self.episodes += len(finished)
            if self.epoch_step_count >= self.epoch_steps:
                self._end_epoch(workers)
            if self._checkpoint_and_stop():
                break

    # -------------------------------------------------------------- helpers
    def _advance(self, n):
        self.steps += n
        self.epoch_step_count += n
        self.steps_since_save += n

    def _end_epoch(self, workers):
        if self.test_environment:
            self._test()
        self.epochs += 1
        now = time.time()
        epoch_time = now - self.last_epoch_time
        logger.store('train/episodes', self.episodes)
        logger.store('train/epochs', self.epochs)
        logger.store('train/seconds', now - self.start_time)
        logger.store('train/epoch_seconds', epoch_time)
        logger.store('train/epoch_steps', self.epoch_step_count)
        logger.store('train/steps', self.steps)
        logger.store('train/worker_steps', self.steps // workers)
        logger.store('train/steps_per_second', self.epoch_step_count / epoch_time)
        self.last_row = logger.dump() if self._writer() else logger.get_current_logger()._row()
        if not self._writer():
            logger.get_current_logger().epoch.clear()
        self.last_epoch_time = time.time()
        self.epoch_step_count = 0

    @staticmethod
    def _writer():
        """Under torchrun the replicas are identical: only rank 0 writes logs / checkpoints."""
        import torch.distributed as dist
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0

    def _checkpoint_and_stop(self):
        stop = self.steps >= self.max_steps
        if (stop or self.steps_since_save >= self.save_steps) and not self._writer():
            self.steps_since_save = self.steps % self.save_steps
        elif stop or self.steps_since_save >= self.save_steps:
            path = os.path.join(logger.get_path(), 'checkpoints')
            if os.path.isdir(path) and self.replace_checkpoint:
                for name in os.listdir(path):
                    if name.startswith('step_'):
                        os.remove(os.path.join(path, name))
            self.agent.save(os.path.join(path, f'step_{self.steps}'))
            self.steps_since_save = self.steps % self.save_steps
        return stop

    def _test(self):
        env = self.test_environment
        if not hasattr(self, 'test_observations'):
            self.test_observations = env.start()
            assert len(self.test_observations) == 1
        for _ in range(self.test_episodes):
            score, length = 0.0, 0
            while True:
                actions = self.agent.test_step(self.test_observations, self.steps)
                host = actions.detach().cpu().numpy() if torch.is_tensor(actions) else actions
                assert not np.isnan(host.sum())
                logger.store('test/action', host, stats=True)
                self.test_observations, infos = env.step(actions)
                self.agent.test_update(**infos, steps=self.steps)
                score += float(infos['rewards'][0])
                length += 1
                if bool(infos['resets'][0]):
                    break
            logger.store('test/episode_score', score, stats=True)
            logger.store('test/episode_length', length, stats=True)
