"""ORACLE / TEST INFRASTRUCTURE -- not part of the product.

numpy definition of the synthetic continuous-control environment
``SynthControl(O, A)`` that stands in for Gym/dm_control physics (SURVEY.md
section 8(d): real MuJoCo physics are not available; BASELINE.json's configs are
"synthetic HalfCheetah-shape env (obs=17, act=6)" etc.).

The same dynamics are implemented by the CUDA kernel
``tonic_b200/csrc/env_step.cu``.  Every arithmetic step is chosen so that numpy
on the CPU and the sm_100a kernel produce bit-identical float32 results:

* the reset state comes from a 32-bit integer hash (murmur3 finaliser) mapped
  exactly to float32 in [-1, 1);
* the dynamics use separately rounded float32 multiplies and adds (the kernel
  uses ``__fmul_rn``/``__fadd_rn`` so no FMA contraction happens);
* the reward is computed from integer-quantised actions/states, summed in
  int64 (order independent), converted once to float32.

The class follows the gym-like protocol the reference expects from an
environment (``/root/reference/tonic/environments/distributed.py:18-50``:
``seed(s)``, ``reset() -> obs``, ``step(a) -> (obs, reward, done, info)``,
``observation_space``/``action_space`` with ``.shape``), so the *unmodified*
reference ``tonic.environments.distribute`` can drive it when golden vectors
are generated (oracle/make_golden.py).
"""

import numpy as np

F32 = np.float32
_M32 = 0xFFFFFFFF

# Constants shared with tonic_b200/csrc/env_step.cu (keep in sync).
DECAY = F32(0.9)
GAIN = F32(0.1)
TERM_LIMIT = F32(1.0)
A_QUANT = F32(1024.0)
X_QUANT = F32(256.0)
A_WEIGHT = 100
X_WEIGHT = 16
COST_SCALE = F32(0.01 * 2.0 ** -20)
ALIVE_BONUS = F32(1.0)


def fmix32(h):
    """murmur3 32-bit finaliser on python ints (exact uint32 arithmetic)."""
    h &= _M32
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & _M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & _M32
    h ^= h >> 16
    return h


def reset_state(seed, episode, size):
    """State of episode number `episode` (0-based) of the env seeded `seed`."""
    k = fmix32((seed & _M32) + 0x9E3779B9 * ((episode + 1) & _M32))
    out = np.empty(size, F32)
    for j in range(size):
        h = fmix32(k ^ ((0x85EBCA6B * (j + 1)) & _M32))
        # 24 random bits -> multiple of 2^-23 in [0, 2) -> exact in float32.
        out[j] = F32(h >> 8) * F32(2.0 ** -23) - F32(1.0)
    return out


def reset_state_vec(seeds, episodes, size):
    """Vectorised `reset_state` for arrays of seeds / episode numbers."""
    seeds = np.asarray(seeds, np.uint64) & np.uint64(_M32)
    episodes = np.asarray(episodes, np.uint64)

    def mix(h):
        h = h & np.uint64(_M32)
        h = h ^ (h >> np.uint64(16))
        h = (h * np.uint64(0x85EBCA6B)) & np.uint64(_M32)
        h = h ^ (h >> np.uint64(13))
        h = (h * np.uint64(0xC2B2AE35)) & np.uint64(_M32)
        h = h ^ (h >> np.uint64(16))
        return h

    k = mix(seeds + np.uint64(0x9E3779B9) * ((episodes + np.uint64(1)) & np.uint64(_M32)))
    j = (np.uint64(0x85EBCA6B) * np.arange(1, size + 1, dtype=np.uint64)) & np.uint64(_M32)
    h = mix(k[:, None] ^ j[None, :])
    return (h >> np.uint64(8)).astype(F32) * F32(2.0 ** -23) - F32(1.0)


def dynamics(x, a):
    """One transition for a batch. x [n,O] f32, a [n,A] (any float dtype).

    Returns (next_x f32 [n,O], reward f32 [n], termination bool [n]).
    """
    x = np.asarray(x, F32)
    a = np.clip(np.asarray(a).astype(F32), F32(-1), F32(1))
    n, size = x.shape
    act = a.shape[1]
    src = a[:, np.arange(size) % act]
    drive = GAIN * src                      # rounded f32 multiply
    nx = DECAY * x                          # rounded f32 multiply
    nx[:, 0] = x[:, 0]                      # coordinate 0 is a pure integrator
    nx = nx + drive                         # rounded f32 add
    qa = np.rint(a * A_QUANT).astype(np.int64)
    qx = np.rint(nx * X_QUANT).astype(np.int64)
    cost = A_WEIGHT * (qa * qa).sum(1) + X_WEIGHT * (qx * qx).sum(1)
    reward = ALIVE_BONUS - cost.astype(F32) * COST_SCALE
    term = np.abs(nx[:, 0]) > TERM_LIMIT
    return nx.astype(F32), reward.astype(F32), term


class Space:
    """Minimal Box-like space (what the reference reads: .shape/.low/.high)."""

    def __init__(self, size):
        self.shape = (size,)
        self.low = -np.ones(size, F32)
        self.high = np.ones(size, F32)
        self.dtype = np.dtype(F32)


class SynthControlEnv:
    """Single synthetic environment with the gym-like protocol."""

    def __init__(self, observation_size=17, action_size=6,
                 max_episode_steps=1000, name=None):
        self.observation_size = observation_size
        self.action_size = action_size
        self.observation_space = Space(observation_size)
        self.action_space = Space(action_size)
        self.max_episode_steps = max_episode_steps
        self._max_episode_steps = max_episode_steps
        self.name = name or f'SynthControl-{observation_size}-{action_size}'
        self._seed = 0
        self._episode = 0
        self.x = np.zeros(observation_size, F32)

    def seed(self, seed):
        self._seed = int(seed)
        self._episode = 0

    def reset(self):
        self.x = reset_state(self._seed, self._episode, self.observation_size)
        self._episode += 1
        return self.x.copy()

    def step(self, action):
        nx, rew, term = dynamics(self.x[None], np.asarray(action)[None])
        self.x = nx[0]
        return self.x.copy(), rew[0], bool(term[0]), {}

    def render(self, *args, **kwargs):
        return None
