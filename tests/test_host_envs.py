"""Host vector environments (tonic_b200/environments/host.py) against the trajectories the
reference's own `Sequential` produced (tests/golden/units.npz, keys env/*), and the forked
`HostParallel` against the in-process `HostSequential`."""

import importlib.util
import os

import numpy as np

from oracle import synth_env  # the numpy environment the golden trajectories were recorded on

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_host():
    """Import host.py without the package __init__ (which loads the CUDA library)."""
    spec = importlib.util.spec_from_file_location(
        'tb_host_envs', os.path.join(ROOT, 'tonic_b200', 'environments', 'host.py'))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def builder():
    return synth_env.SynthControlEnv(7, 3, 6)


def drive(env, actions):
    obs = [env.start()]
    out = dict(next_observations=[], rewards=[], resets=[], terminations=[])
    for a in actions:
        o, infos = env.step(a)
        obs.append(o)
        out['next_observations'].append(infos['observations'])
        for k in ('rewards', 'resets', 'terminations'):
            out[k].append(infos[k])
    return np.array(obs), {k: np.array(v) for k, v in out.items()}


def test_host_sequential_matches_reference(golden):
    host = load_host()
    g = golden('units')
    env = host.distribute_host(builder, 1, 5)
    assert isinstance(env, host.HostSequential) and len(env) == 5 and env.max_episode_steps == 6
    env.initialize(seed=21)
    obs, out = drive(env, g['env/actions'])
    np.testing.assert_array_equal(obs, g['env/observations'])
    for k in out:
        np.testing.assert_array_equal(out[k], g['env/' + k])
        assert out[k].dtype == g['env/' + k].dtype
    assert obs.dtype == np.float32


def test_host_parallel_matches_sequential():
    host = load_host()
    rs = np.random.RandomState(5)
    actions = (rs.normal(size=(25, 6, 3)) * 1.5).astype(np.float32)
    seq = host.distribute_host(builder, 1, 6)
    seq.initialize(seed=3)
    par = host.distribute_host(builder, 3, 2)
    assert isinstance(par, host.HostParallel) and len(par) == 6
    par.initialize(seed=3)
    try:
        a_obs, a_out = drive(seq, actions)
        b_obs, b_out = drive(par, actions)
    finally:
        par.close()
    np.testing.assert_array_equal(a_obs, b_obs)
    for k in a_out:
        np.testing.assert_array_equal(a_out[k], b_out[k])
    assert a_out['resets'].sum() > 0


def load_classic():
    """Import builders.py + classic.py as a stand-alone package (no CUDA library)."""
    import sys
    import types
    pkg = types.ModuleType('tb_envs_pkg')
    pkg.__path__ = [os.path.join(ROOT, 'tonic_b200', 'environments')]
    sys.modules['tb_envs_pkg'] = pkg
    for name in ('builders', 'classic'):
        spec = importlib.util.spec_from_file_location(
            f'tb_envs_pkg.{name}', os.path.join(ROOT, 'tonic_b200', 'environments', f'{name}.py'))
        module = importlib.util.module_from_spec(spec)
        sys.modules[f'tb_envs_pkg.{name}'] = module
        spec.loader.exec_module(module)
    return sys.modules['tb_envs_pkg.classic']


def test_pendulum_on_the_host_grid():
    """`Gym('Pendulum-v1')` (restated dynamics) under the host worker grid: shapes, dtypes,
    action rescaling, the 200-step time-out as a non-terminal reset, determinism per seed."""
    host, classic = load_host(), load_classic()
    env = host.distribute_host(lambda: classic.Gym('Pendulum-v1'), 1, 3)
    assert env.max_episode_steps == 200 and env.observation_space.shape == (3,)
    assert env.action_space.shape == (1,) and float(env.action_space.high[0]) == 1.0
    env.initialize(seed=7)
    first = env.start()
    assert first.shape == (3, 3) and first.dtype == np.float32
    np.testing.assert_allclose(first[:, 0] ** 2 + first[:, 1] ** 2, 1.0, rtol=1e-6)
    rs = np.random.RandomState(0)
    resets = 0
    for t in range(200):
        obs, infos = env.step(rs.uniform(-3, 3, (3, 1)))       # clipped to [-1, 1], scaled to +-2
        assert not infos['terminations'].any()
        assert (infos['rewards'] <= 0).all() and np.abs(infos['observations'][:, 2]).max() <= 8.0
        resets += int(infos['resets'].sum())
    assert resets == 3 and infos['resets'].all()             # all three hit the time limit together
    again = host.distribute_host(lambda: classic.Gym('Pendulum-v1'), 1, 3)
    again.initialize(seed=7)
    np.testing.assert_array_equal(again.start(), first)
    # full torque in one direction for one step from rest at the bottom: theta_dot = 3 * 2 * dt
    task = classic.Gym('Pendulum-v1', time_feature=True)
    task.seed(0)
    task.reset()
    task.environment.state = np.array([0.0, 0.0])
    obs, reward, term, _ = task.step(np.array([5.0]))
    np.testing.assert_allclose(obs[2], 0.3, rtol=1e-6)
    np.testing.assert_allclose(obs[3], -1 + 2 / 200)
    np.testing.assert_allclose(reward, -0.001 * 4.0)
