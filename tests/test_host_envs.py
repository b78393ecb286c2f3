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
