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
    # 391 vector steps of 256 workers per epoch (100,096 steps): the 4th epoch ends at 400,384
    # steps, training stops at the first multiple of 256 >= 410,000 (trainer.py:93-112)
    trainer = train.train(trainer='tonic.Trainer(steps=int(4.1e5), epoch_steps=int(1e5), '
                                  'save_steps=int(1e6), test_episodes=2, show_progress=False)',
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


def _score_curve(seed, fast, iterations=20):
    """Mean score of the episodes finished during every collected segment of one PPO run (2 x 256
    tanh MLPs, 256 environments, 100-step episodes, T = 64, 4 epochs x 8 minibatches of 2048) in
    parity mode (host torch / numpy RNG streams, eager launches) or in the benchmarked fast mode
    (device Philox noise, device Feistel permutations, CUDA-graph replay)."""
    import tonic_b200
    import tonic_b200.torch
    from tonic_b200 import config
    from tonic_b200.utils import logger
    m, n = tonic_b200.torch.models, tonic_b200.torch.normalizers
    config.noise = config.indices = 'device' if fast else 'host'
    logger.store = lambda *a, **k: None
    spec = tonic_b200.environments.SynthControl('HalfCheetah', max_episode_steps=100)
    env = tonic_b200.environments.distribute(lambda: spec, 1, 256)
    env.initialize(seed=seed)
    model = m.ActorCritic(
        actor=m.Actor(encoder=m.ObservationEncoder(), torso=m.MLP((256, 256), torch.nn.Tanh),
                      head=m.DetachedScaleGaussianPolicyHead()),
        critic=m.Critic(encoder=m.ObservationEncoder(), torso=m.MLP((256, 256), torch.nn.Tanh),
                        head=m.ValueHead()),
        observation_normalizer=n.MeanStd())
    agent = tonic_b200.torch.agents.PPO(
        model=model, replay=tonic_b200.replays.Segment(size=64, batch_iterations=4, batch_size=2048))
    agent.initialize(env.observation_space, env.action_space, seed=seed)
    env.start()
    curve = []
    for _ in range(iterations):
        assert agent.rollout(env, 64) == 64
        scores, lengths = env.finished_episodes()
        assert len(scores) > 0
        curve.append(float(np.mean(scores)))
    return np.array(curve)


def test_fast_mode_returns_stay_in_the_parity_mode_band():
    """north_star: "returns matching reference within tolerance" for the BENCHED configuration.
    The parity mode is pinned to the reference step by step (tests/test_gpu_agents.py); here the
    fast mode's learning curves (episode returns, 5 seeds) must be statistically
    indistinguishable from the parity mode's: same start, same improvement, final level within
    the seed-to-seed band."""
    from tonic_b200 import config
    saved = config.noise, config.indices
    try:
        parity = np.array([_score_curve(s, fast=False) for s in range(5)])
        fast = np.array([_score_curve(s, fast=True) for s in range(5)])
    finally:
        config.noise, config.indices = saved
    assert parity.shape == fast.shape == (5, 20)
    print('parity', np.round(parity.mean(0), 2), 'fast', np.round(fast.mean(0), 2))
    p_start, f_start = parity[:, :2].mean(), fast[:, :2].mean()
    p_end, f_end = parity[:, -5:].mean(1), fast[:, -5:].mean(1)
    gain_p, gain_f = p_end.mean() - p_start, f_end.mean() - f_start
    assert gain_p > 1.0 and gain_f > 1.0, (gain_p, gain_f)          # both learn
    # same start (same initial policy and environments, only the noise stream differs)
    assert abs(f_start - p_start) <= 0.25 * gain_p + 0.5, (f_start, p_start)
    band = 3.0 * (p_end.std() + f_end.std()) / np.sqrt(5) + 0.15 * abs(gain_p)
    assert abs(f_end.mean() - p_end.mean()) <= band, (f_end, p_end, band)
    # the mean curve of the fast mode lies inside the parity envelope widened by the same band
    lo, hi = parity.min(0) - band, parity.max(0) + band
    inside = ((fast.mean(0) >= lo) & (fast.mean(0) <= hi)).mean()
    assert inside >= 0.9, (inside, fast.mean(0), lo, hi)
