"""Agent-level parity on the GPU: the tonic_b200 agents, driven through the
reference's protocol, against the golden trajectories / logged losses / final
weights recorded from the unmodified reference (tests/golden/*.npz).

Tolerances: action indices / minibatch indices / env transitions bit-exact;
float32 losses and statistics within 1e-4 relative (BASELINE.json north_star);
weights after the scenario within 2e-4 absolute (they accumulate ~50 Adam steps
of float32 round-off differences)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import scenarios  # noqa: E402  (checker only)
from tests import product  # noqa: E402

ON_POLICY = ['ppo_small', 'ppo_wide', 'ppo_ragged', 'ppo_fullbatch', 'ppo_timefeature', 'ppo_clip',
             'a2c_small']


def by_key(keys, values):
    out = {}
    for k, v in zip(keys, values):
        out.setdefault(str(k), []).append(float(v))
    return out


def check_infos(rec, g, skip=()):
    got = by_key(rec.keys, rec.means)
    ref = by_key(g['info_keys'], g['info_mean'])
    assert set(got) == set(ref), (sorted(got), sorted(ref))
    for k in ref:
        if k in skip:
            continue
        assert len(got[k]) == len(ref[k]), k
        # losses / statistics: 1e-4 relative (+1e-6 absolute for values near zero)
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-4, atol=2e-6, err_msg=k)


def check_weights(agent, g, atol=2e-4):
    w = scenarios.state_arrays(agent.model.state_dict(), 'w/')
    ref_keys = {k.replace('digest_', '') for k in g if k.startswith(('w/', 'digest_w/'))}
    assert ref_keys == set(w)
    for k, v in w.items():
        if k in g:
            np.testing.assert_allclose(v, g[k], rtol=1e-3, atol=atol, err_msg=k)
        else:
            f = v.astype(np.float64).ravel()
            d = g['digest_' + k]
            np.testing.assert_allclose(f[:8], d[2:2 + min(8, f.size)], rtol=1e-3, atol=atol,
                                       err_msg=k)
            np.testing.assert_allclose(np.abs(f).sum(), d[1], rtol=1e-3, err_msg=k)


@pytest.mark.parametrize('name', ON_POLICY)
def test_on_policy_scenario_matches_reference(golden, name):
    g = golden(name)
    cfg = scenarios.SCENARIOS[name]
    rec = scenarios.InfoRecorder()
    agent, env = product.build(cfg, log=rec)
    # initial weights: same torch.nn.Linear init stream as the reference
    w0 = scenarios.state_arrays(agent.model.state_dict(), 'w0/')
    for k, v in w0.items():
        if k in g:
            np.testing.assert_array_equal(v, g[k], err_msg=k)
    actions = product.teacher_forced(agent, env, g, cfg['vector_steps'])
    # sampled actions: same noise stream, float32 MLP round-off only
    # (policy actions come from weights that already took several Adam steps: 3xTF32 /
    # summation-order differences of ~1e-5, cf. the 2e-4 weight tolerance below)
    np.testing.assert_allclose(actions, g['actions'], rtol=1e-5, atol=5e-5)
    check_infos(rec, g)
    check_weights(agent, g)


OFF_POLICY = ['ddpg_small', 'ddpg_ou', 'ddpg_nstep', 'td3_small', 'td3_clip', 'sac_small', 'sac_wrap']


@pytest.mark.parametrize('name', OFF_POLICY)
def test_off_policy_scenario_matches_reference(golden, name):
    g = golden(name)
    cfg = scenarios.SCENARIOS[name]
    rec = scenarios.InfoRecorder()
    agent, env = product.build(cfg, log=rec)
    w0 = scenarios.state_arrays(agent.model.state_dict(), 'w0/')
    for k, v in w0.items():
        if k in g:
            np.testing.assert_array_equal(v, g[k], err_msg=k)
    actions = product.teacher_forced(agent, env, g, cfg['vector_steps'])
    # warm-up actions: the numpy uniform stream (exact up to the float32 cast);
    # afterwards policy actions + numpy / torch noise streams
    # (policy actions come from weights that already took several Adam steps: 3xTF32 /
    # summation-order differences of ~1e-5, cf. the 2e-4 weight tolerance below)
    np.testing.assert_allclose(actions, g['actions'], rtol=1e-5, atol=5e-5)
    check_infos(rec, g)
    check_weights(agent, g)


def test_ppo_iteration_at_the_benched_shape_matches_reference(golden):
    """One PPO iteration at BASELINE.json configs[1]'s shape -- 4096 environments x 128
    steps, 2 x 256 tanh MLPs, 10 epochs x 32 minibatches of 16384 -- against digests of the
    UNMODIFIED reference (tests/golden/ppo_bench.npz, oracle/bench_shape.py).  Host-RNG
    parity mode: torch CPU noise stream, numpy-compatible MT19937 permutations."""
    from oracle import bench_shape
    from tonic_b200 import kernels
    g = golden('ppo_bench')
    cfg = bench_shape.CFG
    rec = scenarios.InfoRecorder()
    agent, env = product.build(cfg, log=rec)
    w0 = bench_shape.weight_digests(agent.model.state_dict(), 'digest_w0/')
    for k, v in w0.items():
        np.testing.assert_allclose(v, g[k], rtol=1e-12, atol=0, err_msg=k)     # same init stream

    class HostEnv:      # numpy protocol on top of the device environment
        def start(self):
            return env.start(host=True)

        def step(self, actions):
            return env.step(np.asarray(actions, np.float32))
    out = bench_shape.drive(agent, HostEnv(), cfg)
    # environment trajectory: bit-exact -> identical float64 digests and counts
    for k in ('observation_digest', 'reward_digest', 'reset_count', 'termination_count'):
        np.testing.assert_array_equal(out[k], g[k], err_msg=k)
    assert g['termination_count'].sum() > 0
    # sampled actions: same eps stream, float32 round-off of the 2 x 256 MLP
    np.testing.assert_allclose(out['action_sample'], g['action_sample'], rtol=1e-5, atol=5e-5)
    np.testing.assert_allclose(out['action_digest'], g['action_digest'], rtol=1e-5)
    # Logged statistics.  First epoch (32 minibatches per network, identical weights up to the
    # round-off of <= 32 Adam steps): the 1e-4 relative bar of the small scenarios.  Whole update
    # (320 sequential Adam steps per network, 3xTF32 vs the reference's fp32 BLAS summation
    # order): the drift of the weights shows up in the statistics; the surrogate loss is a mean of
    # (normalised advantage x ratio) terms of unit scale that cancel to ~1e-2, so its error is
    # absolute (1e-4 of the terms' scale), not relative to the cancelled value.
    got = by_key(rec.keys, rec.means)
    ref = by_key(g['info_keys'], g['info_mean'])
    assert set(got) == set(ref) and all(len(got[k]) == len(ref[k]) for k in ref)
    assert len(got['critic/loss']) == 320 and len(got['actor/loss']) == int(got['actor/iterations'][0])
    # First epoch: 1e-4 relative, with absolute floors where the statistic is a cancelling mean
    # (the surrogate loss: unit-scale terms cancelling to ~1e-2), a difference of log-probabilities
    # (KL ~1e-3) or a count (one of 16384 samples crossing the clip boundary = 6.1e-5).
    # Whole update: the reference ITSELF is only reproducible to the spread recorded in
    # tests/golden/ppo_bench_thread_sensitivity.json (same torch code, 1 vs 8 BLAS threads): the
    # bar is 3x that spread (+ the first-epoch bar).
    import json
    import os
    spread = json.load(open(os.path.join(os.path.dirname(__file__), 'golden',
                                         'ppo_bench_thread_sensitivity.json')))['whole_update_max_abs']
    floor_first = {'actor/loss': 1e-4, 'actor/kl': 2e-6, 'actor/clip_fraction': 3.1 / 16384}
    worst = {}
    for k in ref:
        a, b = np.array(got[k]), np.array(ref[k])
        if k in ('actor/stop', 'actor/iterations', 'critic/iterations'):
            np.testing.assert_array_equal(a, b, err_msg=k)
            continue
        err = np.abs(a - b)
        worst[k] = (float(err[:32].max()), float((err[:32] / (np.abs(b[:32]) + 1e-12)).max()),
                    float(err.max()), float((err / (np.abs(b) + 1e-12)).max()))
        bar_first = 1e-4 * np.abs(b) + floor_first.get(k, 2e-6)
        assert (err[:32] <= bar_first[:32]).all(), (k, worst[k])
        assert (err <= bar_first + 3.0 * spread[k]).all(), (k, worst[k], spread[k])
    print('benched shape (abs, rel) errors first epoch | whole update:', worst)
    # final weights (320 Adam steps per network at B = 16384)
    w = bench_shape.weight_digests(agent.model.state_dict(), 'digest_w/')
    assert set(w) == {k for k in g if k.startswith('digest_w/')}
    sizes = {'digest_w/' + k: v.numel() for k, v in agent.model.state_dict().items()}
    for k, v in w.items():
        np.testing.assert_allclose(v[2:], g[k][2:], rtol=1e-3, atol=2e-4, err_msg=k)
        # sum of |w|: 1e-3 relative; tiny tensors (log_scale: 6 values near zero) get the per-element
        # absolute tolerance of the line above
        np.testing.assert_allclose(v[1], g[k][1], rtol=1e-3, atol=2e-4 * min(sizes[k], 8), err_msg=k)


def test_greedy_and_test_time_actions_match_reference_kats():
    """SURVEY.md 8(c) KAT2 / KAT5 on the GPU (a29b): default PPO / DDPG, seed 0, O=17, A=6:
    first stochastic step, the DDPG warm-up uniform actions and the greedy test-time action."""
    import tonic_b200
    import tonic_b200.torch

    class Space:
        def __init__(self, n):
            self.shape = (n,)
    ppo = tonic_b200.torch.agents.PPO()
    ppo.initialize(Space(17), Space(6), seed=0)
    actions = ppo.step(np.zeros((2, 17), np.float32), 0)
    np.testing.assert_allclose(actions[0, :3], [-0.7685214, -0.05233827, -0.05380505], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.asarray(ppo.last_log_probs.cpu()), [-5.1127973, -5.57454], rtol=1e-5)
    # the PPO / A2C test-time action is a fresh stochastic sample (a2c.py:87-90): next
    # draw of the same torch CPU stream
    ddpg = tonic_b200.torch.agents.DDPG()
    ddpg.initialize(Space(17), Space(6), seed=0)
    warm = ddpg.step(np.zeros((2, 17), np.float32), 0)
    np.testing.assert_allclose(warm[0, :3], [0.09762701, 0.43037873, 0.20552675], rtol=1e-6)
    greedy = ddpg.test_step(np.zeros((1, 17), np.float32), 0)
    np.testing.assert_allclose(np.asarray(greedy)[0, :3], [0.0708199, 0.0229248, 0.05138565],
                               rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('name', ['ppo_wide', 'ddpg_small', 'sac_small'])
def test_test_time_actions_match_oracle(golden, name):
    """a29b: after a teacher-forced scenario the product's `test_step` (stochastic sample for
    PPO, greedy / tanh(mean) for DDPG / SAC; a2c.py:87-90, ddpg.py:78-81, sac.py:48-51) agrees
    with the oracle port that ran the same scenario on the CPU."""
    import torch
    from oracle import port
    g = golden(name)
    cfg = scenarios.SCENARIOS[name]
    agent, env = product.build(cfg)
    product.teacher_forced(agent, env, g, cfg['vector_steps'])
    oagent, oenv = port.build(cfg)
    scenarios.drive(oagent, oenv, cfg['vector_steps'])
    obs = np.asarray(g['observations'][-1][:1], np.float32)
    torch.manual_seed(1234)
    ours = np.asarray(agent.test_step(obs, 0))
    torch.manual_seed(1234)
    theirs = np.asarray(oagent.test_step(obs, 0))
    np.testing.assert_allclose(ours, theirs, rtol=1e-4, atol=5e-5)


def test_cuda_graph_sections_match_eager():
    """Fast mode (device noise + device permutations): replaying the captured rollout /
    update graphs gives the same weights as issuing the kernels one by one."""
    import torch
    from tonic_b200 import config
    cfg = dict(scenarios.SCENARIOS['ppo_wide'], workers=64,
               segment=dict(size=8, batch_iterations=2, batch_size=128))
    results = []
    old = config.noise, config.indices, config.graphs, config.fused_rollout
    try:
        for use_graphs in (False, True):
            config.noise, config.indices, config.graphs = 'device', 'device', use_graphs
            config.fused_rollout = False        # the per-step launches are what gets captured
            agent, env = product.build(cfg)
            env.start()
            for _ in range(5):          # 2 eager warm-ups, capture, 2 replays
                assert agent.rollout(env, cfg['segment']['size']) == cfg['segment']['size']
            torch.cuda.synchronize()
            results.append(torch.cat([n.params for n in agent.model.networks()]).cpu())
            if use_graphs:
                assert agent._update_graph.graph is not None
                assert agent._rollout_graph.graph is not None
    finally:
        config.noise, config.indices, config.graphs, config.fused_rollout = old
    # same kernels, same device-resident RNG streams -> bit-identical parameters
    assert torch.equal(results[0], results[1])


@pytest.mark.parametrize('kind', ['PPO', 'A2C'])
def test_fused_step_kernel_matches_separate_kernels(kind):
    """One vector step as actor forward + act_env_step_kernel (sample, log-prob, normaliser record,
    environment step in one launch) against the chain gauss_sample -> counter_add ->
    moments_record -> env_step: same Philox positions and arithmetic -> bit-identical segments,
    environment state and episode log; normaliser sums to double round-off."""
    import torch
    from tonic_b200 import config
    base = scenarios.SCENARIOS['ppo_wide' if kind == 'PPO' else 'a2c_small']
    seg = dict(base['segment'], size=24)
    cfg = dict(base, workers=100, max_episode_steps=9, segment=seg)     # ragged tile + many resets
    old = config.noise, config.indices, config.graphs, config.fused_rollout, config.fused_step
    out = []
    try:
        for fused in (False, True):
            config.noise, config.indices, config.graphs = 'device', 'device', False
            config.fused_rollout, config.fused_step = False, fused
            agent, env = product.build(cfg)
            env.start()
            norm = agent.model.observation_normalizer
            sums = []
            record = norm.update
            norm.update = lambda: (sums.append(norm.sums.clone()), record())[1]
            # 10 single steps through the stepwise entry, then the rest of the segment
            assert agent.rollout(env, 10) == 10
            assert agent.rollout(env, seg['size']) == seg['size'] - 10
            torch.cuda.synchronize()
            n_ep = int(env.episode_count.item())
            out.append(dict(
                seg={k: v.clone() for k, v in agent.replay.buffers.items()
                     if k in ('observations', 'actions', 'next_observations', 'rewards', 'resets',
                              'terminations', 'log_probs')},
                sums=sums[0], state=env.state.clone(), obs=env.observations.clone(), episodes=n_ep,
                scores=torch.sort(env.episode_scores[:n_ep])[0].clone(),
                counter=int(agent._noise_counter.item())))
    finally:
        config.noise, config.indices, config.graphs, config.fused_rollout, config.fused_step = old
    a, b = out
    for k in a['seg']:
        assert torch.equal(a['seg'][k], b['seg'][k]), k
    assert float(a['seg']['resets'].sum()) > 100       # the reset branch was exercised
    assert torch.allclose(a['sums'], b['sums'], rtol=1e-12, atol=1e-9)
    assert a['sums'][-1].item() == b['sums'][-1].item() == 24 * 100
    assert torch.equal(a['state'], b['state']) and torch.equal(a['obs'], b['obs'])
    assert a['episodes'] == b['episodes'] > 0 and torch.equal(a['scores'], b['scores'])
    assert a['counter'] == b['counter'] == 24 * 100


@pytest.mark.parametrize('kind', ['PPO', 'A2C'])
def test_fused_rollout_matches_per_step_launches(kind):
    """The persistent rollout kernel (one launch per segment) against the per-step chain
    actor forward -> gauss_sample -> env_step -> moments_record on the FFMA path: same
    Philox streams, same tile arithmetic -> bit-identical segments, environment state and
    episode log; normaliser sums agree to double round-off (different summation order)."""
    import torch
    from tonic_b200 import config
    base = scenarios.SCENARIOS['ppo_wide' if kind == 'PPO' else 'a2c_small']
    seg = dict(base['segment'], size=24)
    cfg = dict(base, workers=100, max_episode_steps=9, segment=seg)     # ragged tile + many resets
    old = config.noise, config.indices, config.graphs, config.fused_rollout, config.gemm
    out = []
    try:
        for fused in (False, True):
            config.noise, config.indices, config.graphs = 'device', 'device', False
            config.fused_rollout, config.gemm = fused, 'ffma'
            agent, env = product.build(cfg)
            env.start()
            norm = agent.model.observation_normalizer
            sums = []
            record = norm.update
            norm.update = lambda: (sums.append(norm.sums.clone()), record())[1]
            assert agent.rollout(env, seg['size']) == seg['size']
            torch.cuda.synchronize()
            first = {k: v.clone() for k, v in agent.replay.buffers.items()
                     if k in ('observations', 'actions', 'next_observations', 'rewards',
                              'resets', 'terminations', 'log_probs')}
            n_ep = int(env.episode_count.item())
            log = (torch.sort(env.episode_scores[:n_ep])[0].clone(),
                   torch.sort(env.episode_lengths[:n_ep])[0].clone(), env.state.clone())
            for _ in range(2):
                agent.rollout(env, seg['size'])
            torch.cuda.synchronize()
            out.append(dict(
                first=first, sums=sums[0], state=env.state.clone(), obs=env.observations.clone(),
                episodes=n_ep, scores=log[0], lengths=log[1], first_state=log[2],
                params=torch.cat([n.params for n in agent.model.networks()]).cpu()))
    finally:
        config.noise, config.indices, config.graphs, config.fused_rollout, config.gemm = old
    a, b = out
    for k in a['first']:
        assert torch.equal(a['first'][k], b['first'][k]), k
    assert float(a['first']['resets'].sum()) > 100       # the reset branch was exercised
    assert torch.allclose(a['sums'], b['sums'], rtol=1e-12, atol=1e-9)
    assert a['episodes'] == b['episodes'] and a['episodes'] > 0
    torch.testing.assert_close(a['params'], b['params'], rtol=0, atol=1e-5)
    # after three segments the environments are still on identical trajectories unless a
    # 1-ulp difference of the normaliser moved an action; compare loosely
    torch.testing.assert_close(a['state'], b['state'], rtol=0, atol=1e-3)
    assert torch.equal(a['lengths'], b['lengths']) and torch.equal(a['scores'], b['scores'])
    assert torch.equal(a['first_state'], b['first_state'])


@pytest.mark.parametrize('kind', ['DDPG', 'TD3', 'SAC'])
def test_off_policy_fast_path_graph_replay_matches_eager(kind):
    """Fast mode of the off-policy agents (device Philox noise, device-drawn replay indices, ring
    state on the device): `rollout()` -- act -> environment -> ring store -> record -> update per
    vector step -- replayed as CUDA graphs gives bit-identical parameters and replay contents to
    issuing the same kernels one by one; warm-up (uniform actions), collection without updates
    and collection with updates are all exercised."""
    import torch
    from tonic_b200 import config
    cfg = dict(agent=kind, obs=11, act=3, workers=64, max_episode_steps=13, seed=4, hidden=(256, 256),
               start_steps=64 * 4,
               buffer=dict(size=64 * 40, batch_iterations=3, batch_size=32,
                           steps_before_batches=64 * 8, steps_between_batches=50))
    old = config.noise, config.indices, config.graphs
    out = []
    try:
        for use_graphs in (False, True):
            config.noise, config.indices, config.graphs = 'device', 'device', use_graphs
            agent, env = product.build(cfg)
            env.start()
            assert agent.can_rollout(env)
            steps = 0
            for _ in range(9):                       # 4 vector steps per call
                assert agent.rollout(env, 4, steps=steps) == 4
                steps += 4 * 64
            torch.cuda.synchronize()
            if use_graphs:
                assert any(sec.graph is not None for sec in agent._sections.values())
            params = torch.cat([n.params for n in agent.model.networks()]).cpu()
            ring = {k: v.clone().cpu() for k, v in agent.replay.buffers.items()}
            out.append((params, ring, agent.replay.index, agent.replay.size,
                        agent.replay._ring.cpu().tolist()))
    finally:
        config.noise, config.indices, config.graphs = old
    (p0, r0, i0, s0, d0), (p1, r1, i1, s1, d1) = out
    assert torch.isfinite(p0).all() and (i0, s0, d0) == (i1, s1, d1)
    assert d0 == [36 % 40, 36, 36 * 64]              # device ring state = host mirror
    for k in r0:
        assert torch.equal(torch.nan_to_num(r0[k]), torch.nan_to_num(r1[k])), k
    assert torch.equal(p0, p1)
    # the agents did train: online and target networks differ from each other
    nets = agent.model.networks()
    assert not torch.equal(nets[0].params, nets[-1].params)


def test_host_protocol_fast_path_matches_plain_path():
    """The reference protocol with numpy arrays (agent.step / environment.step / agent.update)
    in the product configuration: every call is one captured graph over pinned staging buffers
    (kernels.HostBridge).  Same kernels, same device RNG streams as the call-by-call path ->
    identical actions, observations, segment contents and parameters after the updates."""
    import torch
    from tonic_b200 import config
    cfg = dict(scenarios.SCENARIOS['ppo_wide'], workers=48, max_episode_steps=7,
               segment=dict(size=6, batch_iterations=2, batch_size=96))
    old = config.noise, config.indices, config.graphs
    out = []
    try:
        for use_graphs in (False, True):
            config.noise, config.indices, config.graphs = 'device', 'device', use_graphs
            agent, env = product.build(cfg)
            obs = env.start(host=True)
            trace, steps = [], 0
            for t in range(20):                      # three updates (T = 6) + a partial segment
                actions = agent.step(obs, steps)
                assert isinstance(actions, np.ndarray) and actions.dtype == np.float32
                obs, infos = env.step(actions)
                assert infos['resets'].dtype == np.bool_ and obs.dtype == np.float32
                agent.update(**infos, steps=steps)
                steps += 48
                trace.append((actions.copy(), obs.copy(), infos['rewards'].copy(), infos['resets'].copy()))
            torch.cuda.synchronize()
            seg = {k: v.clone().cpu() for k, v in agent.replay.buffers.items()}
            params = torch.cat([n.params for n in agent.model.networks()]).cpu()
            out.append((trace, seg, params, agent.replay.index))
            if use_graphs:
                assert agent._host_sections[('step', 48)][0].graph is not None
                assert env._host_section.graph is not None
    finally:
        config.noise, config.indices, config.graphs = old
    (t0, s0, p0, i0), (t1, s1, p1, i1) = out
    assert i0 == i1 == 2
    for a, b in zip(t0, t1):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
    assert sum(int(r[3].sum()) for r in t0) > 0          # resets happened
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k
    assert torch.equal(p0, p1)
