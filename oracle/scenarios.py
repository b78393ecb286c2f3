"""ORACLE / TEST INFRASTRUCTURE -- not part of the product.

Scenario configurations and the backend-agnostic driver loop used to
  * generate golden vectors from the UNMODIFIED reference (make_golden.py),
  * pin the oracle port against those vectors (tests/test_oracle_*.py),
  * check the CUDA product against both (tests/test_gpu_*.py).

A scenario is a plain dict; each backend (reference / oracle port / product)
has a small adapter that turns it into (agent, environment) objects that obey
the reference's duck-typed protocol
(/root/reference/tonic/agents/agent.py:7-34 and
/root/reference/tonic/environments/distributed.py:18-58).

The driver issues exactly the call sequence of the reference training loop
(/root/reference/tonic/utils/trainer.py:42-55): `agent.step(obs, steps)` ->
`environment.step(actions)` -> `agent.update(**infos, steps=steps)` with
`steps` being the count BEFORE the increment by the number of workers.
"""

import numpy as np

# name -> config.  Sizes are small so the oracle finishes in seconds and the
# fixtures stay small; max_episode_steps is short to exercise time-outs, and
# the integrator coordinate of SynthControl produces true terminations.
SCENARIOS = {
    'ppo_small': dict(
        agent='PPO', obs=17, act=6, workers=8, max_episode_steps=11, seed=0,
        hidden=(64, 64), vector_steps=40,
        segment=dict(size=16, batch_iterations=3, batch_size=32)),
    'ppo_wide': dict(
        agent='PPO', obs=17, act=6, workers=16, max_episode_steps=25, seed=3,
        hidden=(256, 256), vector_steps=34,
        segment=dict(size=16, batch_iterations=4, batch_size=64)),
    'ppo_ragged': dict(   # last minibatch of every epoch is short (120 % 32)
        agent='PPO', obs=5, act=2, workers=6, max_episode_steps=7, seed=11,
        hidden=(64, 64), vector_steps=45,
        segment=dict(size=20, batch_iterations=2, batch_size=32)),
    'ppo_fullbatch': dict(  # batch_size=None: same full batch every iteration
        agent='PPO', obs=17, act=6, workers=4, max_episode_steps=9, seed=5,
        hidden=(64, 64), vector_steps=26,
        segment=dict(size=12, batch_iterations=5, batch_size=None)),
    'ppo_timefeature': dict(   # build_environment(time_feature=True): wrappers.py:25-54
        agent='PPO', obs=5, act=2, workers=6, max_episode_steps=9, seed=21, time_feature=True,
        hidden=(64, 64), vector_steps=36,
        segment=dict(size=12, batch_iterations=3, batch_size=24)),
    'ppo_clip': dict(     # gradient_clip > 0 in both updaters (actors.py:96-98, critics.py:24-25)
        agent='PPO', obs=9, act=3, workers=8, max_episode_steps=11, seed=7, gradient_clip=0.05,
        hidden=(256, 256), vector_steps=34,
        segment=dict(size=16, batch_iterations=3, batch_size=32)),
    'a2c_small': dict(
        agent='A2C', obs=17, act=6, workers=8, max_episode_steps=11, seed=1,
        hidden=(64, 64), vector_steps=40,
        segment=dict(size=16, batch_iterations=3, batch_size=32)),
    'ddpg_small': dict(
        agent='DDPG', obs=11, act=3, workers=4, max_episode_steps=13, seed=2,
        hidden=(256, 256), vector_steps=40, start_steps=60,
        buffer=dict(size=400, batch_iterations=4, batch_size=16,
                    steps_before_batches=48, steps_between_batches=16)),
    'ddpg_ou': dict(      # OrnsteinUhlenbeckActionNoise (explorations/noisy.py:53-88), SURVEY 8f rank 2
        agent='DDPG', obs=9, act=2, workers=4, max_episode_steps=7, seed=5, exploration='ou',
        hidden=(64, 64), vector_steps=40, start_steps=40,
        buffer=dict(size=400, batch_iterations=3, batch_size=16,
                    steps_before_batches=48, steps_between_batches=16)),
    'ddpg_nstep': dict(   # Buffer(return_steps=3): n-step accumulation (replays/buffers.py:58-79)
        agent='DDPG', obs=6, act=2, workers=4, max_episode_steps=5, seed=12,
        hidden=(64, 64), vector_steps=40, start_steps=60,
        buffer=dict(size=400, return_steps=3, batch_iterations=3, batch_size=16,
                    steps_before_batches=48, steps_between_batches=16)),
    'td3_small': dict(
        agent='TD3', obs=11, act=3, workers=4, max_episode_steps=13, seed=4,
        hidden=(256, 256), vector_steps=40, start_steps=60,
        buffer=dict(size=400, batch_iterations=4, batch_size=16,
                    steps_before_batches=48, steps_between_batches=16)),
    'td3_clip': dict(     # joint-norm clip over both critics (critics.py:177-178) + actor clip
        agent='TD3', obs=7, act=2, workers=4, max_episode_steps=9, seed=9, gradient_clip=0.1,
        hidden=(64, 64), vector_steps=40, start_steps=60,
        buffer=dict(size=400, batch_iterations=4, batch_size=16,
                    steps_before_batches=48, steps_between_batches=16)),
    'sac_small': dict(
        agent='SAC', obs=11, act=3, workers=4, max_episode_steps=13, seed=6,
        hidden=(256, 256), vector_steps=40, start_steps=60,
        buffer=dict(size=400, batch_iterations=4, batch_size=16,
                    steps_before_batches=48, steps_between_batches=16)),
    'sac_wrap': dict(   # ring buffer wraps around (max_size = 40 // 4 = 10 rows)
        agent='SAC', obs=7, act=2, workers=4, max_episode_steps=9, seed=8,
        hidden=(64, 64), vector_steps=36, start_steps=40,
        buffer=dict(size=40, batch_iterations=3, batch_size=8,
                    steps_before_batches=24, steps_between_batches=8)),
}


class InfoRecorder:
    """Collects every `logger.store(key, value)` an agent makes."""

    def __init__(self):
        self.keys, self.means, self.abss = [], [], []

    def __call__(self, key, value, stats=False):
        if hasattr(value, 'detach'):
            value = value.detach().cpu().numpy()
        v = np.atleast_1d(np.asarray(value, np.float64))
        self.keys.append(key)
        self.means.append(v.mean())
        self.abss.append(np.abs(v).mean())

    def arrays(self):
        return dict(info_keys=np.array(self.keys),
                    info_mean=np.array(self.means, np.float64),
                    info_abs=np.array(self.abss, np.float64))


def _host(x):
    if hasattr(x, 'detach'):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def drive(agent, environment, vector_steps):
    """Runs the reference call sequence and returns the recorded trajectory."""
    observations = environment.start()
    workers = len(observations)
    rec = dict(start_observations=_host(observations).copy(),
               actions=[], observations=[], next_observations=[], rewards=[],
               resets=[], terminations=[])
    steps = 0
    for _ in range(vector_steps):
        actions = agent.step(observations, steps)
        rec['actions'].append(_host(actions).astype(np.float64))
        observations, infos = environment.step(actions)
        rec['observations'].append(_host(observations).copy())
        rec['next_observations'].append(_host(infos['observations']).copy())
        rec['rewards'].append(_host(infos['rewards']).copy())
        rec['resets'].append(_host(infos['resets']).astype(np.bool_))
        rec['terminations'].append(_host(infos['terminations']).astype(np.bool_))
        agent.update(**infos, steps=steps)
        steps += workers
    return {k: np.array(v) for k, v in rec.items()}


def state_arrays(state_dict, prefix='w/'):
    return {prefix + k: _host(v).astype(np.float32)
            for k, v in state_dict.items()}
