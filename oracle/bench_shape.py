"""ORACLE / TEST INFRASTRUCTURE -- not part of the product.

Parity at the BENCHED shape (BASELINE.json configs[1]): one PPO iteration of
4096 environments x 128 vector steps, 2 x 256 tanh MLPs, E = 10 epochs of 32
minibatches of 16384 transitions, driven through the reference call sequence
(/root/reference/tonic/utils/trainer.py:42-55).

A full teacher-forcing trajectory at this size would be ~100 MB, so the fixture
is made small in two ways:

* the actions handed to `environment.step` are NOT the agent's samples but a
  closed-form table `driving_actions(t)` (integer hash -> float32) that every
  backend regenerates; the agent still runs `step` on every observation and its
  own sampled actions / log-probs are what `agent.update` stores in the segment
  (a2c.py:48-50,58-64), so the update consumes exactly what the reference's
  would.  The environment trajectory is then identical on every backend
  (bit-exact environment, checked separately);
* large arrays are recorded as digests (float64 sum, sum of squares) plus a
  strided sample.

Used by oracle/make_golden.py (unmodified reference -> tests/golden/ppo_bench.npz),
tests/test_oracle_golden.py (oracle port, CPU, opt-in because it takes minutes)
and tests/test_gpu_agents.py (CUDA product).
"""

import numpy as np

from . import scenarios

CFG = dict(
    agent='PPO', obs=17, act=6, workers=4096, max_episode_steps=1000, seed=0,
    hidden=(256, 256), vector_steps=128,
    segment=dict(size=128, batch_iterations=10, batch_size=16384))

SAMPLE_STRIDE = 128        # workers kept in the strided samples (4096 / 128 = 32 per step)


def driving_actions(t, workers, act):
    """float32 [workers, act] in [-1.25, 1.25): murmur3 finaliser of the flat index,
    24 bits -> exact float32 (the clip of ActionRescaler is exercised)."""
    i = (np.arange(workers * act, dtype=np.uint64) + np.uint64(t * workers * act)) \
        & np.uint64(0xFFFFFFFF)
    h = i * np.uint64(0x9E3779B1) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = h * np.uint64(0x85EBCA6B) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = h * np.uint64(0xC2B2AE35) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    u = (h >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -23) - np.float32(1.0)
    return (u * np.float32(1.25)).reshape(workers, act)


def digest(x):
    f = np.asarray(x, np.float64).ravel()
    return np.array([f.sum(), np.square(f).sum()])


def drive(agent, environment, cfg=CFG, to_host=None):
    """The reference call sequence with the driving-action table; returns the record."""
    to_host = to_host or scenarios._host
    observations = environment.start()
    workers, act = cfg['workers'], cfg['act']
    rec = dict(action_digest=[], action_sample=[], observation_digest=[], reward_digest=[],
               reset_count=[], termination_count=[])
    steps = 0
    for t in range(cfg['vector_steps']):
        actions = to_host(agent.step(observations, steps))
        rec['action_digest'].append(digest(actions))
        rec['action_sample'].append(np.asarray(actions, np.float64)[::SAMPLE_STRIDE].copy())
        observations, infos = environment.step(driving_actions(t, workers, act))
        rec['observation_digest'].append(digest(to_host(observations)))
        rec['reward_digest'].append(digest(to_host(infos['rewards'])))
        rec['reset_count'].append(int(np.asarray(to_host(infos['resets'])).sum()))
        rec['termination_count'].append(int(np.asarray(to_host(infos['terminations'])).sum()))
        agent.update(**infos, steps=steps)
        steps += workers
    return {k: np.array(v) for k, v in rec.items()}


def weight_digests(state_dict, prefix):
    out = {}
    for k, v in state_dict.items():
        f = scenarios._host(v).astype(np.float64).ravel()
        out[prefix + k] = np.concatenate([[f.sum(), np.abs(f).sum()], f[:8],
                                          np.zeros(max(0, 8 - f.size))])
    return out
