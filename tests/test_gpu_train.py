"""End-to-end: the reference-style command line trains PPO with the fused trainer on
the device environment; the logged return improves, checkpoints are written with the
reference's state_dict keys and can be resumed."""

import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_cli_learns_and_checkpoints(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    from tonic_b200 import config, train
    monkeypatch.setattr(config, 'noise', 'device')
    monkeypatch.setattr(config, 'indices', 'device')
    common = dict(
        header='import tonic_b200.torch',
        agent="tonic.torch.agents.PPO(replay=tonic.replays.Segment(size=64, batch_iterations=4, "
              "batch_size=2048))",
        environment="tonic.environments.SynthControl('HalfCheetah', max_episode_steps=100)",
        test_environment=None, before_training=None, after_training=None, parallel=1,
        sequential=256, seed=0, name='ppo-test', environment_name='synth', checkpoint='last')
    trainer = train.train(trainer='tonic.Trainer(steps=int(4e5), epoch_steps=int(1e5), '
                                  'save_steps=int(4e5), test_episodes=2, show_progress=False)',
                          path=None, **common)
    run = os.path.join('synth', 'ppo-test', '0')
    rows = open(os.path.join(run, 'log.csv')).read().strip().splitlines()
    header = rows[0].split(',')
    table = np.array([[float(c) if c != 'None' else np.nan for c in r.split(',')]
                      for r in rows[1:]])
    assert len(table) == 4
    score = table[:, header.index('train/episode_score/mean')]
    assert score[-1] > score[0] + 5, score            # learning happens
    for key in ('train/steps_per_second', 'train/action/mean', 'test/episode_score/mean',
                'actor/loss', 'critic/loss', 'actor/kl', 'train/episodes'):
        assert key in header, key
    saved = sorted(os.listdir(os.path.join(run, 'checkpoints')))
    assert len(saved) == 1 and saved[0].startswith('step_4'), saved
    state = torch.load(os.path.join(run, 'checkpoints', saved[0]))
    assert 'actor.torso.model.0.weight' in state and 'critic.head.v_layer.bias' in state
    assert 'critic.encoder.observation_normalizer._mean' in state
    # resume from the checkpoint (train.py:22-75): weights are restored
    common2 = dict(common, agent=None, environment=None, header=None)
    trainer2 = train.train(trainer='tonic.Trainer(steps=int(1e5), epoch_steps=int(1e5), '
                                   'save_steps=int(1e6), test_episodes=1, show_progress=False)',
                           path=run, **common2)
    assert trainer2.steps >= int(1e5)
