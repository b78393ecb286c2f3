"""ORACLE / TEST INFRASTRUCTURE -- not part of the product.

Generates the golden fixtures under tests/golden/ by importing and running the
UNMODIFIED reference from /root/reference (through the test-only `gym` /
`termcolor` import shims in oracle/shims).  /root/reference only exists in the
build container, so the fixtures are committed; nothing at test/bench time
reads /root/reference.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Recorded with torch 2.11.0 (CPU) / numpy 2.3.5.
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'shims'))
sys.path.insert(1, '/root/reference')
sys.path.insert(2, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import gym  # noqa: E402  (the shim)
import tonic  # noqa: E402  (the reference)
import tonic.torch  # noqa: E402

from oracle import bench_shape, scenarios, synth_env  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


# --------------------------------------------------------------------------
# Reference adapter: scenario config -> reference objects.
# --------------------------------------------------------------------------

def reference_environment(cfg):
    def raw(name):
        env = synth_env.SynthControlEnv(
            cfg['obs'], cfg['act'], cfg['max_episode_steps'], name=name)
        high = np.ones(cfg['act'], np.float32)
        env.action_space = gym.spaces.Box(-high, high)
        if cfg.get('time_feature'):      # TimeFeature reads observation_space.low / high / dtype
            bound = np.full(cfg['obs'], np.inf, np.float32)
            env.observation_space = gym.spaces.Box(-bound, bound)
        return gym.wrappers.TimeLimit(env, cfg['max_episode_steps'])

    def builder():
        return tonic.environments.builders.build_environment(
            raw, 'synth', time_feature=bool(cfg.get('time_feature', False)))

    env = tonic.environments.distribute(builder, 1, cfg['workers'])
    env.initialize(seed=cfg['seed'])
    return env


def reference_agent(cfg):
    m, n, u = tonic.torch.models, tonic.torch.normalizers, tonic.torch.updaters
    hidden = tuple(cfg['hidden'])
    kind = cfg['agent']
    if kind in ('PPO', 'A2C'):
        model = m.ActorCritic(
            actor=m.Actor(
                encoder=m.ObservationEncoder(),
                torso=m.MLP(hidden, torch.nn.Tanh),
                head=m.DetachedScaleGaussianPolicyHead()),
            critic=m.Critic(
                encoder=m.ObservationEncoder(),
                torso=m.MLP(hidden, torch.nn.Tanh),
                head=m.ValueHead()),
            observation_normalizer=n.MeanStd())
        replay = tonic.replays.Segment(**cfg['segment'])
        cls = getattr(tonic.torch.agents, kind)
        clip = cfg.get('gradient_clip', 0)
        if clip:
            actor_updater = (u.ClippedRatio if kind == 'PPO' else u.StochasticPolicyGradient)(
                gradient_clip=clip)
            return cls(model=model, replay=replay, actor_updater=actor_updater,
                       critic_updater=u.VRegression(gradient_clip=clip))
        return cls(model=model, replay=replay)

    critic = m.Critic(
        encoder=m.ObservationActionEncoder(),
        torso=m.MLP(hidden, torch.nn.ReLU), head=m.ValueHead())
    if kind == 'SAC':
        head = m.GaussianPolicyHead(
            loc_activation=torch.nn.Identity,
            distribution=m.SquashedMultivariateNormalDiag)
    else:
        head = m.DeterministicPolicyHead()
    actor = m.Actor(
        encoder=m.ObservationEncoder(),
        torso=m.MLP(hidden, torch.nn.ReLU), head=head)
    wrapper = m.ActorCriticWithTargets if kind == 'DDPG' else \
        m.ActorTwinCriticWithTargets
    model = wrapper(actor=actor, critic=critic,
                    observation_normalizer=n.MeanStd())
    replay = tonic.replays.Buffer(**cfg['buffer'])
    if kind == 'SAC':
        exploration = tonic.explorations.NoActionNoise(cfg['start_steps'])
    elif cfg.get('exploration') == 'ou':
        exploration = tonic.explorations.OrnsteinUhlenbeckActionNoise(
            start_steps=cfg['start_steps'])
    else:
        exploration = tonic.explorations.NormalActionNoise(
            start_steps=cfg['start_steps'])
    cls = getattr(tonic.torch.agents, kind)
    clip = cfg.get('gradient_clip', 0)
    if clip:
        actor_cls = dict(DDPG=u.DeterministicPolicyGradient, TD3=u.DeterministicPolicyGradient,
                         SAC=u.TwinCriticSoftDeterministicPolicyGradient)[kind]
        critic_cls = dict(DDPG=u.DeterministicQLearning, TD3=u.TwinCriticDeterministicQLearning,
                          SAC=u.TwinCriticSoftQLearning)[kind]
        return cls(model=model, replay=replay, exploration=exploration,
                   actor_updater=actor_cls(gradient_clip=clip),
                   critic_updater=critic_cls(gradient_clip=clip))
    return cls(model=model, replay=replay, exploration=exploration)


def run_reference_scenario(name):
    cfg = scenarios.SCENARIOS[name]
    env = reference_environment(cfg)
    agent = reference_agent(cfg)
    agent.initialize(env.observation_space, env.action_space, seed=cfg['seed'])
    out = scenarios.state_arrays(agent.model.state_dict(), 'w0/')
    rec = scenarios.InfoRecorder()
    tonic.utils.logger.store = rec      # capture what the agent logs
    out.update(scenarios.drive(agent, env, cfg['vector_steps']))
    out.update(rec.arrays())
    out.update(scenarios.state_arrays(agent.model.state_dict(), 'w/'))
    return out


# --------------------------------------------------------------------------
# Unit-level known-answer vectors.
# --------------------------------------------------------------------------

def unit_vectors():
    out = {}
    # lambda_returns (replays/utils.py:4-19) on random data with resets.
    for tag, (T, N, seed) in dict(a=(4, 2, 0), b=(16, 8, 1), c=(33, 5, 2)).items():
        rs = np.random.RandomState(seed)
        v = rs.normal(size=(T, N)).astype(np.float32)
        nv = rs.normal(size=(T, N)).astype(np.float32)
        r = rs.normal(size=(T, N)).astype(np.float32)
        resets = (rs.uniform(size=(T, N)) < 0.2).astype(np.float32)
        terms = (resets * (rs.uniform(size=(T, N)) < 0.5)).astype(np.float32)
        ret = tonic.replays.lambda_returns(v, nv, r, resets, terms, 0.99, 0.97)
        for k, a in dict(values=v, next_values=nv, rewards=r, resets=resets,
                         terminations=terms, returns=ret).items():
            out[f'lam_{tag}/{k}'] = a
    # KAT1 of SURVEY.md section 8(c).
    rs = np.random.RandomState(0)
    v, nv, r = (rs.normal(size=(4, 2)).astype(np.float32) for _ in range(3))
    resets = np.zeros((4, 2), np.float32)
    terms = np.zeros((4, 2), np.float32)
    resets[1, 0] = resets[2, 1] = 1
    terms[2, 1] = 1
    out['kat1/returns'] = tonic.replays.lambda_returns(
        v, nv, r, resets, terms, 0.99, 0.97)

    # Segment minibatch index streams (replays/segments.py:50-65).
    for tag, (size, workers, iters, bs, seed) in dict(
            a=(16, 8, 3, 32, 0), b=(20, 6, 2, 32, 11), c=(128, 64, 2, 2048, 7)
    ).items():
        seg = tonic.replays.Segment(size=size, batch_iterations=iters, batch_size=bs)
        seg.initialize(seed)
        for t in range(size):
            seg.store(ids=np.arange(workers) + t * workers)
        batches = [b['ids'].astype(np.int64) for b in seg.get('ids')]
        out[f'segidx_{tag}/indices'] = np.concatenate(batches)
        out[f'segidx_{tag}/lengths'] = np.array([len(b) for b in batches])
        out[f'segidx_{tag}/cfg'] = np.array([size, workers, iters, bs, seed])

    # Advantage normalisation in get_full (replays/segments.py:41-46).
    rs = np.random.RandomState(5)
    seg = tonic.replays.Segment(size=32, batch_iterations=1, batch_size=None)
    seg.initialize(0)
    for t in range(32):
        seg.store(rewards=rs.normal(size=16), resets=rs.uniform(size=16) < .1,
                  terminations=np.zeros(16))
    vals = rs.normal(size=32 * 16).astype(np.float32)
    nvals = rs.normal(size=32 * 16).astype(np.float32)
    seg.compute_returns(vals, nvals)
    full = seg.get_full('advantages', 'returns')
    out['adv/values'] = vals
    out['adv/next_values'] = nvals
    out['adv/rewards'] = seg.buffers['rewards']
    out['adv/resets'] = seg.buffers['resets']
    out['adv/advantages'] = full['advantages']
    out['adv/returns'] = full['returns']

    # Buffer sampling stream (replays/buffers.py:81-91) incl. wrap-around.
    buf = tonic.replays.Buffer(size=40, batch_iterations=3, batch_size=8,
                               steps_before_batches=0, steps_between_batches=1)
    buf.initialize(9)
    for t in range(14):          # max_size = 40 // 4 = 10 rows -> wraps
        buf.store(ids=np.arange(4) + 4 * t, terminations=np.arange(4) == t % 4)
    got = [b for b in buf.get('ids', 'discounts', steps=100)]
    out['bufidx/ids'] = np.stack([b['ids'] for b in got])
    out['bufidx/discounts'] = np.stack([b['discounts'] for b in got])

    # MeanStd record/update (torch/normalizers/mean_stds.py:44-70).
    ms = tonic.torch.normalizers.MeanStd()
    ms.initialize((5,))
    rs = np.random.RandomState(2)
    batches = [rs.normal(1.5, 2.0, size=(7, 5)).astype(np.float32) for _ in range(6)]
    snaps = []
    for i, b in enumerate(batches):
        ms.record(b)
        if i % 2 == 1:
            ms.update()
            snaps.append(np.stack([ms._mean.numpy().copy(), ms._std.numpy().copy()]))
    out['meanstd/batches'] = np.stack(batches)
    out['meanstd/snapshots'] = np.stack(snaps)

    # numpy legacy RandomState streams the host side reproduces natively.
    out['kat3/shuffle8'] = np.random.RandomState(0).permutation(8) * 0
    a = np.arange(8)
    np.random.RandomState(0).shuffle(a)
    out['kat3/shuffle8'] = a
    a = np.arange(1000)
    rs = np.random.RandomState(123)
    rs.shuffle(a)
    out['kat3/shuffle1000_seed123'] = a.copy()
    rs.shuffle(a)
    out['kat3/shuffle1000_seed123_second'] = a.copy()
    out['kat4/randint1000_5'] = np.random.RandomState(0).randint(1000, size=5)
    rs = np.random.RandomState(77)
    out['kat4/randint_70000_64'] = rs.randint(70000, size=64)
    out['kat4/randint_5e9_16'] = rs.randint(5 * 10 ** 9, size=16)
    rs = np.random.RandomState(42)
    out['kat6/uniform'] = rs.uniform(-1, 1, (3, 4))
    out['kat6/normal'] = rs.normal(size=(5, 3))
    out['kat6/normal_after'] = rs.normal(size=(2, 3))

    # Synthetic env dynamics under the reference Sequential (distributed.py:28-58).
    cfg = dict(obs=7, act=3, workers=5, max_episode_steps=6, seed=21)
    env = reference_environment(cfg)
    obs = [env.start()]
    rs = np.random.RandomState(3)
    acts, nobs, rew, res, ter = [], [], [], [], []
    for t in range(30):
        a = (rs.normal(size=(5, 3)) * 1.5).astype(np.float32)
        o, infos = env.step(a)
        acts.append(a); obs.append(o); nobs.append(infos['observations'])
        rew.append(infos['rewards']); res.append(infos['resets'])
        ter.append(infos['terminations'])
    out['env/actions'] = np.array(acts)
    out['env/observations'] = np.array(obs)
    out['env/next_observations'] = np.array(nobs)
    out['env/rewards'] = np.array(rew)
    out['env/resets'] = np.array(res)
    out['env/terminations'] = np.array(ter)
    return out


def run_reference_bench_shape():
    """One PPO iteration of the unmodified reference at the benched shape
    (oracle/bench_shape.py): digests + strided samples."""
    import time
    cfg = bench_shape.CFG
    t0 = time.time()
    env = reference_environment(cfg)
    agent = reference_agent(cfg)
    agent.initialize(env.observation_space, env.action_space, seed=cfg['seed'])
    out = bench_shape.weight_digests(agent.model.state_dict(), 'digest_w0/')
    rec = scenarios.InfoRecorder()
    tonic.utils.logger.store = rec
    out.update(bench_shape.drive(agent, env, cfg))
    out.update(rec.arrays())
    out.update(bench_shape.weight_digests(agent.model.state_dict(), 'digest_w/'))
    out['torch_threads'] = np.array([torch.get_num_threads()])
    print(f'ppo_bench: {time.time() - t0:.1f} s,',
          sum(k == 'critic/loss' for k in rec.keys), 'critic updates,',
          sum(k == 'actor/loss' for k in rec.keys), 'actor updates')
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    only = [a for a in sys.argv[1:]]
    if 'ppo_bench' in only:
        np.savez_compressed(os.path.join(OUT, 'ppo_bench.npz'), **run_reference_bench_shape())
        only.remove('ppo_bench')
        if not only:
            return
    if not only:
        np.savez_compressed(os.path.join(OUT, 'units.npz'), **unit_vectors())
        print('units.npz')
    for name in (only or scenarios.SCENARIOS):
        data = run_reference_scenario(name)
        if max(scenarios.SCENARIOS[name]['hidden']) > 64:
            # keep wide-model fixtures small: final weights as float16-free
            # digests (sum, abs-sum, first 8 entries) instead of full tensors.
            slim = {}
            for k, v in data.items():
                if k.startswith('w/') or k.startswith('w0/'):
                    f = v.astype(np.float64).ravel()
                    slim['digest_' + k] = np.concatenate(
                        [[f.sum(), np.abs(f).sum()], f[:8], np.zeros(max(0, 8 - f.size))])
                else:
                    slim[k] = v
            data = slim
        np.savez_compressed(os.path.join(OUT, f'{name}.npz'), **data)
        n_up = sum(k.endswith('critic/loss') for k in data['info_keys'])
        print(name, 'critic updates logged:', n_up,
              'resets:', int(data['resets'].sum()),
              'terminations:', int(data['terminations'].sum()))


if __name__ == '__main__':
    main()
