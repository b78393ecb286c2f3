"""Scenario config (oracle/scenarios.py) -> product objects (tonic_b200), the
third adapter next to the reference one (oracle/make_golden.py) and the oracle
port (oracle/port.py::build)."""

import numpy as np
import torch


def build(cfg, log=None):
    import tonic_b200
    import tonic_b200.torch
    from tonic_b200.utils import logger
    m, n = tonic_b200.torch.models, tonic_b200.torch.normalizers
    if log is not None:
        logger.store = log
    spec = tonic_b200.environments.SynthControl(
        'synth', cfg['obs'], cfg['act'], cfg['max_episode_steps'],
        time_feature=cfg.get('time_feature', False))
    env = tonic_b200.environments.distribute(lambda: spec, 1, cfg['workers'])
    env.initialize(seed=cfg['seed'])
    hidden = tuple(cfg['hidden'])
    kind = cfg['agent']
    if kind in ('PPO', 'A2C'):
        model = m.ActorCritic(
            actor=m.Actor(encoder=m.ObservationEncoder(), torso=m.MLP(hidden, torch.nn.Tanh),
                          head=m.DetachedScaleGaussianPolicyHead()),
            critic=m.Critic(encoder=m.ObservationEncoder(), torso=m.MLP(hidden, torch.nn.Tanh),
                            head=m.ValueHead()),
            observation_normalizer=n.MeanStd())
        replay = tonic_b200.replays.Segment(**cfg['segment'])
        extra = {}
        clip = cfg.get('gradient_clip', 0)
        if clip:
            u = tonic_b200.torch.updaters
            extra = dict(
                actor_updater=(u.ClippedRatio if kind == 'PPO' else u.StochasticPolicyGradient)(
                    gradient_clip=clip),
                critic_updater=u.VRegression(gradient_clip=clip))
        agent = getattr(tonic_b200.torch.agents, kind)(model=model, replay=replay, **extra)
    else:
        critic = m.Critic(encoder=m.ObservationActionEncoder(),
                          torso=m.MLP(hidden, torch.nn.ReLU), head=m.ValueHead())
        if kind == 'SAC':
            head = m.GaussianPolicyHead(loc_activation=torch.nn.Identity,
                                        distribution=m.SquashedMultivariateNormalDiag)
        else:
            head = m.DeterministicPolicyHead()
        actor = m.Actor(encoder=m.ObservationEncoder(), torso=m.MLP(hidden, torch.nn.ReLU),
                        head=head)
        wrapper = m.ActorCriticWithTargets if kind == 'DDPG' else m.ActorTwinCriticWithTargets
        model = wrapper(actor=actor, critic=critic, observation_normalizer=n.MeanStd())
        replay = tonic_b200.replays.Buffer(**cfg['buffer'])
        if kind == 'SAC':
            exploration = tonic_b200.explorations.NoActionNoise(cfg['start_steps'])
        elif cfg.get('exploration') == 'ou':
            exploration = tonic_b200.explorations.OrnsteinUhlenbeckActionNoise(
                start_steps=cfg['start_steps'])
        else:
            exploration = tonic_b200.explorations.NormalActionNoise(
                start_steps=cfg['start_steps'])
        extra = {}
        clip = cfg.get('gradient_clip', 0)
        if clip:
            u = tonic_b200.torch.updaters
            actor_cls = dict(DDPG=u.DeterministicPolicyGradient, TD3=u.DeterministicPolicyGradient,
                             SAC=u.TwinCriticSoftDeterministicPolicyGradient)[kind]
            critic_cls = dict(DDPG=u.DeterministicQLearning, TD3=u.TwinCriticDeterministicQLearning,
                              SAC=u.TwinCriticSoftQLearning)[kind]
            extra = dict(actor_updater=actor_cls(gradient_clip=clip),
                         critic_updater=critic_cls(gradient_clip=clip))
        agent = getattr(tonic_b200.torch.agents, kind)(
            model=model, replay=replay, exploration=exploration, **extra)
    agent.initialize(env.observation_space, env.action_space, seed=cfg['seed'])
    return agent, env


def teacher_forced(agent, env, golden, vector_steps, rows=slice(None)):
    """Drives agent + device env with the golden trajectory's inputs: the env
    receives the reference's actions (so its outputs must be bit-identical), the
    agent sees the reference's observations (its actions must agree to float32
    round-off)."""
    obs = env.start(host=True)
    np.testing.assert_array_equal(obs, golden['start_observations'][rows])
    workers = golden['start_observations'].shape[0]     # GLOBAL workers (trainer.py:54)
    steps = 0
    actions = []
    for t in range(vector_steps):
        a = agent.step(obs, steps)
        actions.append(np.asarray(a, np.float64))
        ref_a = golden['actions'][t][rows]
        ref_a = ref_a.astype(np.float32) if np.asarray(a).dtype == np.float32 else ref_a
        obs, infos = env.step(ref_a)
        np.testing.assert_array_equal(obs, golden['observations'][t][rows])
        np.testing.assert_array_equal(infos['observations'], golden['next_observations'][t][rows])
        np.testing.assert_array_equal(infos['rewards'], golden['rewards'][t][rows])
        np.testing.assert_array_equal(infos['resets'], golden['resets'][t][rows])
        np.testing.assert_array_equal(infos['terminations'], golden['terminations'][t][rows])
        agent.update(**infos, steps=steps)
        steps += workers
    return np.array(actions)
