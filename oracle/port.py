"""ORACLE / TEST INFRASTRUCTURE -- not part of the product.

CPU restatement (numpy + torch-CPU) of the reference's data-parallel hot path:
the synchronous vector-env collector, the Segment / Buffer replays, lambda
returns, the running observation normaliser, and the PPO / A2C / DDPG / TD3 /
SAC update loops.  Every function cites the reference file:line it follows
(paths relative to /root/reference/tonic).  The arithmetic that the reference
delegates to third-party libraries (torch 2.11.0: nn.Linear, tanh/relu/
softplus, distributions.Normal, MSELoss, optim.Adam, autograd; numpy 2.3.5:
RandomState.{shuffle,randint,normal,uniform}, ndarray.std/mean) is delegated
to the same installed libraries here, exactly at the reference's call sites.

PINNING: tests/test_oracle_golden.py checks this port against the fixtures in
tests/golden/*.npz, which were produced by running the unmodified reference
(oracle/make_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
(tonic_b200/) never does.
"""

import numpy as np
import torch
import torch.nn.functional as F

from . import synth_env

F32 = np.float32


# ---------------------------------------------------------------------------
# Vector environment  (environments/distributed.py:8-58, Sequential)
# ---------------------------------------------------------------------------

class VectorEnv:
    """M environments stepped in sequence inside one process."""

    def __init__(self, obs, act, workers, max_episode_steps, first_worker=0, time_feature=False):
        self.envs = [synth_env.SynthControlEnv(obs, act, max_episode_steps)
                     for _ in range(workers)]
        self.max_episode_steps = max_episode_steps
        self.observation_space = self.envs[0].observation_space
        if time_feature:       # environments/wrappers.py:25-54 (low = -1, high = 1)
            self.observation_space = synth_env.Space(obs + 1)
        self.time_feature = time_feature
        self.action_space = self.envs[0].action_space
        self.first_worker = first_worker

    def _timed(self, ob, steps):
        """TimeFeature.reset / step (wrappers.py:40-54): append low + (high - low) * steps / max."""
        if not self.time_feature:
            return ob
        v = -1 + (1 - -1) * (steps / self.max_episode_steps) if steps else -1
        return np.append(ob, v)

    def initialize(self, seed):
        # distributed.py:18-20: env i is seeded seed + i.
        for i, env in enumerate(self.envs):
            env.seed(seed + self.first_worker + i)

    def start(self):
        # distributed.py:22-26
        self.lengths = np.zeros(len(self.envs), int)
        return np.array([self._timed(env.reset(), 0) for env in self.envs], F32)

    def step(self, actions):
        # distributed.py:28-58.  ActionRescaler with unit bounds
        # (environments/wrappers.py:18-22) is clip(a, -1, 1); the synthetic
        # env clips internally.
        acting_obs, trans_obs, rews, resets, terms = [], [], [], [], []
        for i, env in enumerate(self.envs):
            ob, rew, term, _ = env.step(actions[i])
            self.lengths[i] += 1
            ob = self._timed(ob, self.lengths[i])
            # distributed.py:39-40: a time-out resets but is not a termination.
            reset = term or self.lengths[i] == self.max_episode_steps
            trans_obs.append(ob)
            rews.append(rew)
            resets.append(reset)
            terms.append(term)
            if reset:
                ob = self._timed(env.reset(), 0)
                self.lengths[i] = 0
            acting_obs.append(ob)
        infos = dict(observations=np.array(trans_obs, F32),
                     rewards=np.array(rews, F32),
                     resets=np.array(resets, np.bool_),
                     terminations=np.array(terms, np.bool_))
        return np.array(acting_obs, F32), infos


class ParallelVectorEnv:
    """P forked worker groups of M sequential environments each, synchronous scatter / gather
    over a pipe per group and one result queue (environments/distributed.py:69-155): worker
    j = g * M + i is seeded seed + j, actions are split group-major (`np.split`, :137)."""

    def __init__(self, obs, act, groups, per_group, max_episode_steps):
        self.shape = (obs, act, max_episode_steps)
        self.groups, self.per_group = groups, per_group
        probe = VectorEnv(obs, act, 1, max_episode_steps)
        self.observation_space, self.action_space = probe.observation_space, probe.action_space
        self.max_episode_steps = max_episode_steps

    def initialize(self, seed):
        import multiprocessing
        ctx = multiprocessing.get_context('fork')
        self.queue = ctx.Queue()
        self.pipes, self.procs = [], []

        def proc(pipe, index, group_seed):
            envs = VectorEnv(*self.shape[:2], self.per_group, self.shape[2])
            envs.initialize(group_seed)
            self.queue.put((index, envs.start()))
            while True:
                actions = pipe.recv()
                if actions is None:
                    return
                self.queue.put((index, envs.step(actions)))

        for i in range(self.groups):
            pipe, worker_end = ctx.Pipe()
            self.pipes.append(pipe)
            p = ctx.Process(target=proc, args=(worker_end, i, seed + i * self.per_group))
            p.daemon = True
            p.start()
            self.procs.append(p)

    def start(self):
        parts = [None] * self.groups
        for _ in range(self.groups):
            index, observations = self.queue.get()
            parts[index] = observations
        return np.concatenate(parts)

    def step(self, actions):
        for part, pipe in zip(np.split(np.asarray(actions), self.groups), self.pipes):
            pipe.send(part)
        outs = [None] * self.groups
        for _ in range(self.groups):
            index, out = self.queue.get()
            outs[index] = out
        observations = np.concatenate([o[0] for o in outs])
        infos = {k: np.concatenate([o[1][k] for o in outs])
                 for k in ('observations', 'rewards', 'resets', 'terminations')}
        return observations, infos

    def close(self):
        for pipe in self.pipes:
            pipe.send(None)
        for p in self.procs:
            p.join(timeout=5)


# ---------------------------------------------------------------------------
# Replays
# ---------------------------------------------------------------------------

def lambda_returns(values, next_values, rewards, resets, terminations,
                   discount_factor, trace_decay):
    """replays/utils.py:4-19 -- reverse scan over the time axis, float32."""
    out = np.zeros_like(values)
    carry = next_values[-1]
    for t in range(len(rewards) - 1, -1, -1):
        boot = (1 - trace_decay) * next_values[t] + trace_decay * carry
        boot *= (1 - resets[t])
        boot += resets[t] * next_values[t]
        boot *= (1 - terminations[t])
        out[t] = carry = rewards[t] + discount_factor * boot
    return out


def normalized_advantages(returns, values):
    """replays/segments.py:41-46 -- whole-array mean/std (ddof=0)."""
    adv = returns - values
    std = adv.std()
    if std != 0:
        adv = (adv - adv.mean()) / std
    return adv


class SegmentStore:
    """replays/segments.py:6-78."""

    def __init__(self, size=4096, batch_iterations=80, batch_size=None,
                 discount_factor=0.99, trace_decay=0.97):
        self.size, self.iterations, self.batch_size = size, batch_iterations, batch_size
        self.gamma, self.lam = discount_factor, trace_decay

    def initialize(self, seed=None):
        self.rng = np.random.RandomState(seed)      # segments.py:20
        self.data, self.index = None, 0

    def ready(self):
        return self.index == self.size               # segments.py:24-25

    def store(self, **kw):                           # segments.py:27-36
        if self.data is None:
            self.workers = len(next(iter(kw.values())))
            self.data = {k: np.zeros((self.size,) + np.array(v).shape, F32)
                         for k, v in kw.items()}
        for k, v in kw.items():
            self.data[k][self.index] = v
        self.index += 1

    def flat(self, *keys):                           # segments.py:38-48
        self.index = 0
        if 'advantages' in keys:
            self.data['advantages'] = normalized_advantages(
                self.data['returns'], self.data['values'])
        return {k: self.data[k].reshape((-1,) + self.data[k].shape[2:])
                for k in keys}

    def batches(self, *keys):                        # segments.py:50-65
        full = self.flat(*keys)
        if self.batch_size is None:
            for _ in range(self.iterations):
                yield full
            return
        total = self.size * self.workers
        order = np.arange(total)
        for _ in range(self.iterations):
            self.rng.shuffle(order)
            for lo in range(0, total, self.batch_size):
                idx = order[lo:lo + self.batch_size]
                yield {k: v[idx] for k, v in full.items()}

    def compute_returns(self, values, next_values):  # segments.py:67-78
        shape = self.data['rewards'].shape
        self.data['values'] = values.reshape(shape)
        self.data['next_values'] = next_values.reshape(shape)
        self.data['returns'] = lambda_returns(
            self.data['values'], self.data['next_values'], self.data['rewards'],
            self.data['resets'], self.data['terminations'], self.gamma, self.lam)


class RingStore:
    """replays/buffers.py:4-91 (n-step accumulation: :58-79)."""

    def __init__(self, size=int(1e6), return_steps=1, batch_iterations=50,
                 batch_size=100, discount_factor=0.99,
                 steps_before_batches=int(1e4), steps_between_batches=50):
        self.return_steps = return_steps
        self.full_size, self.iterations, self.batch_size = size, batch_iterations, batch_size
        self.gamma = discount_factor
        self.before, self.between = steps_before_batches, steps_between_batches

    def initialize(self, seed=None):                 # buffers.py:21-26
        self.rng = np.random.RandomState(seed)
        self.data, self.index, self.count, self.last_steps = None, 0, 0, 0

    def ready(self, steps):                          # buffers.py:28-31
        return steps >= self.before and (steps - self.last_steps) >= self.between

    def store(self, **kw):                           # buffers.py:33-56
        if 'terminations' in kw:
            kw['discounts'] = F32(1 - kw['terminations']) * self.gamma
        if self.data is None:
            self.workers = len(next(iter(kw.values())))
            self.rows = self.full_size // self.workers
            self.data = {k: np.full((self.rows,) + np.array(v).shape, np.nan, F32)
                         for k, v in kw.items()}
        for k, v in kw.items():
            self.data[k][self.index] = v
        if self.return_steps > 1:                    # buffers.py:58-79
            rewards, next_obs, discounts = kw['rewards'], kw['next_observations'], kw['discounts']
            masks = np.ones(self.workers, F32)
            for i in range(min(self.count, self.return_steps - 1)):
                index = (self.index - i - 1) % self.rows
                masks *= (1 - self.data['resets'][index])
                new_rewards = self.data['rewards'][index] + self.data['discounts'][index] * rewards
                self.data['rewards'][index] = (1 - masks) * self.data['rewards'][index] + masks * new_rewards
                new_discounts = self.data['discounts'][index] * discounts
                self.data['discounts'][index] = ((1 - masks) * self.data['discounts'][index]
                                                 + masks * new_discounts)
                self.data['next_observations'][index] = (
                    (1 - masks)[:, None] * self.data['next_observations'][index]
                    + masks[:, None] * next_obs)
        self.index = (self.index + 1) % self.rows
        self.count = min(self.count + 1, self.rows)

    def batches(self, *keys, steps):                 # buffers.py:81-91
        for _ in range(self.iterations):
            idx = self.rng.randint(self.count * self.workers, size=self.batch_size)
            rows, cols = idx // self.workers, idx % self.workers
            yield {k: self.data[k][rows, cols] for k in keys}
        self.last_steps = steps


# ---------------------------------------------------------------------------
# Observation normaliser  (torch/normalizers/mean_stds.py:5-74)
# ---------------------------------------------------------------------------

class RunningMoments:
    def __init__(self, size):
        self.mean = np.zeros(size, F32)
        self.std = np.ones(size, F32)
        self.mean_sq = np.square(self.mean)
        self.count = 0
        self.new_sum, self.new_sum_sq, self.new_count = 0, 0, 0
        self.t_mean = torch.zeros(size)
        self.t_std = torch.ones(size)

    def __call__(self, x):                           # mean_stds.py:34-39
        return (x - self.t_mean) / self.t_std

    def record(self, values):                        # mean_stds.py:44-48
        for v in values:                             # sequential f32 sums
            self.new_sum += v
            self.new_sum_sq += np.square(v)
            self.new_count += 1

    def update(self):                                # mean_stds.py:50-74
        total = self.count + self.new_count
        w_old, w_new = self.count / total, self.new_count / total
        self.mean = w_old * self.mean + w_new * (self.new_sum / self.new_count)
        self.mean_sq = w_old * self.mean_sq + w_new * (self.new_sum_sq / self.new_count)
        var = np.maximum(self.mean_sq - np.square(self.mean), 0)
        self.std = np.maximum(np.sqrt(var), 1e-2)
        self.count = total
        self.new_sum, self.new_sum_sq, self.new_count = 0, 0, 0
        self.t_mean = torch.as_tensor(self.mean, dtype=torch.float32).clone()
        self.t_std = torch.as_tensor(self.std, dtype=torch.float32).clone()


# ---------------------------------------------------------------------------
# Networks: 2-layer MLP torso + head, parameters named as the reference's
# state_dict (torch/models/utils.py:4-23, actors.py, critics.py).
# ---------------------------------------------------------------------------

class Net:
    """Ordered name -> Parameter map; layers are created with torch.nn.Linear
    so the default initialisation and the global-RNG consumption match the
    reference's creation order (models/utils.py:11-19)."""

    def __init__(self, prefix):
        self.prefix = prefix
        self.p = {}

    def linear(self, name, fan_in, fan_out):
        layer = torch.nn.Linear(fan_in, fan_out)
        self.p[name + '.weight'] = layer.weight
        self.p[name + '.bias'] = layer.bias

    def torso(self, fan_in, hidden):
        for i, width in enumerate(hidden):
            self.linear(f'torso.model.{2 * i}', fan_in, width)
            fan_in = width
        return fan_in

    def run_torso(self, x, act):
        i = 0
        while f'torso.model.{2 * i}.weight' in self.p:
            x = act(F.linear(x, self.p[f'torso.model.{2 * i}.weight'],
                             self.p[f'torso.model.{2 * i}.bias']))
            i += 1
        return x

    def parameters(self):
        return list(self.p.values())

    def named(self):
        return {self.prefix + k: v for k, v in self.p.items()}

    def copy_from(self, other):                      # actor_critics.py:64-66
        for k in self.p:
            self.p[k].data.copy_(other.p[k].data)

    def freeze(self):
        for v in self.p.values():
            v.requires_grad = False


def gaussian_actor(prefix, obs, act, hidden):
    """Actor + DetachedScaleGaussianPolicyHead (models/actors.py:37-66,118-137)."""
    net = Net(prefix)
    width = net.torso(obs, hidden)
    net.linear('head.loc_layer.0', width, act)
    net.p['head.log_scale'] = torch.nn.Parameter(torch.zeros(1, act))
    return net


def gaussian_actor_forward(net, observations):
    # Quirk a17: the actor's encoder has NO observation normaliser
    # (models/actors.py:128-129 passes it positionally into `action_space`).
    h = net.run_torso(observations, torch.tanh)
    loc = torch.tanh(F.linear(h, net.p['head.loc_layer.0.weight'],
                              net.p['head.loc_layer.0.bias']))
    scale = F.softplus(net.p['head.log_scale']) + 1e-8   # actors.py:63
    scale = torch.clamp(scale, 1e-4, 1.).repeat(observations.shape[0], 1)
    return torch.distributions.normal.Normal(loc, scale)


def value_critic(prefix, inputs, hidden):
    """Critic + ValueHead (models/critics.py:4-20,70-90)."""
    net = Net(prefix)
    width = net.torso(inputs, hidden)
    net.linear('head.v_layer', width, 1)
    return net


def value_critic_forward(net, norm, act_fn, observations, actions=None):
    x = norm(observations)          # critics DO normalise (critics.py:81-83)
    if actions is not None:         # encoders.py:28-31
        x = torch.cat([x, actions], dim=-1)
    h = net.run_torso(x, act_fn)
    out = F.linear(h, net.p['head.v_layer.weight'], net.p['head.v_layer.bias'])
    return torch.squeeze(out, -1)


def deterministic_actor(prefix, obs, act, hidden):
    """Actor + DeterministicPolicyHead (models/actors.py:101-115)."""
    net = Net(prefix)
    width = net.torso(obs, hidden)
    net.linear('head.action_layer.0', width, act)
    return net


def deterministic_actor_forward(net, observations):
    h = net.run_torso(observations, torch.relu)
    return torch.tanh(F.linear(h, net.p['head.action_layer.0.weight'],
                               net.p['head.action_layer.0.bias']))


def squashed_actor(prefix, obs, act, hidden):
    """Actor + GaussianPolicyHead(loc Identity) (models/actors.py:69-98)."""
    net = Net(prefix)
    width = net.torso(obs, hidden)
    net.linear('head.loc_layer.0', width, act)
    net.linear('head.scale_layer.0', width, act)
    return net


def squashed_actor_forward(net, observations):
    h = net.run_torso(observations, torch.relu)
    loc = F.linear(h, net.p['head.loc_layer.0.weight'], net.p['head.loc_layer.0.bias'])
    scale = F.softplus(F.linear(h, net.p['head.scale_layer.0.weight'],
                                net.p['head.scale_layer.0.bias']))
    scale = torch.clamp(scale, 1e-4, 1)
    return torch.distributions.normal.Normal(loc, scale)


def squashed_rsample_with_log_prob(dist):
    """SquashedMultivariateNormalDiag.rsample_with_log_prob (actors.py:11-16)."""
    raw = dist.rsample()
    squashed = torch.tanh(raw)
    log_probs = dist.log_prob(raw)
    log_probs -= torch.log(1 - squashed ** 2 + 1e-6)
    return squashed, log_probs


def as_f32(x):
    return torch.as_tensor(x, dtype=torch.float32)


# ---------------------------------------------------------------------------
# Agents
# ---------------------------------------------------------------------------

class OnPolicyOracle:
    """A2C (torch/agents/a2c.py:20-127) and PPO (torch/agents/ppo.py:7-67)."""

    def __init__(self, kind, hidden, segment, log=None, actor_lr=3e-4,
                 critic_lr=1e-3, ratio_clip=0.2, kl_threshold=0.015,
                 entropy_coeff=0.0, gradient_clip=0):
        self.gradient_clip = gradient_clip            # actors.py:37-38,96-98; critics.py:24-25
        self.kind, self.hidden = kind, tuple(hidden)
        self.replay = SegmentStore(**segment)
        self.log = log or (lambda *a, **k: None)
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.ratio_clip, self.kl_threshold = ratio_clip, kl_threshold
        self.entropy_coeff = entropy_coeff

    def initialize(self, observation_space, action_space, seed=None):
        if seed is not None:                          # torch/agents/agent.py:11-15
            np.random.seed(seed)
            import random
            random.seed(seed)
            torch.manual_seed(seed)
        obs, act = observation_space.shape[0], action_space.shape[0]
        self.norm = RunningMoments(obs)
        self.actor = gaussian_actor('actor.', obs, act, self.hidden)
        self.critic = value_critic('critic.', obs, self.hidden)
        self.replay.initialize(seed)
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), lr=self.actor_lr)
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=self.critic_lr)

    def state_dict(self):
        out = dict(self.actor.named())
        out.update(self.critic.named())
        for pre in ('critic.encoder.observation_normalizer.', 'observation_normalizer.'):
            out[pre + '_mean'] = self.norm.t_mean
            out[pre + '_std'] = self.norm.t_std
        return out

    def step(self, observations, steps):              # a2c.py:41-52,75-85
        with torch.no_grad():
            dist = gaussian_actor_forward(self.actor, as_f32(observations))
            actions = dist.sample()
            log_probs = dist.log_prob(actions).sum(dim=-1)
        self.last_observations = np.array(observations, copy=True)
        self.last_actions = actions.numpy().copy()
        self.last_log_probs = log_probs.numpy().copy()
        return actions.numpy()

    def test_step(self, observations, steps):         # a2c.py:87-90 (stochastic)
        with torch.no_grad():
            return gaussian_actor_forward(self.actor, as_f32(observations)).sample().numpy()

    def update(self, observations, rewards, resets, terminations, steps):
        # a2c.py:58-73
        self.replay.store(
            observations=self.last_observations, actions=self.last_actions,
            next_observations=observations, rewards=rewards, resets=resets,
            terminations=terminations, log_probs=self.last_log_probs)
        self.norm.record(self.last_observations)
        if self.replay.ready():
            self._update()

    def _values(self, observations):
        return value_critic_forward(self.critic, self.norm, torch.tanh, observations)

    def _evaluate(self):                              # a2c.py:92-99
        batch = self.replay.flat('observations', 'next_observations')
        with torch.no_grad():
            values = self._values(as_f32(batch['observations']))
            next_values = self._values(as_f32(batch['next_observations']))
        self.replay.compute_returns(values.numpy(), next_values.numpy())

    def _critic_step(self, observations, returns):    # updaters/critics.py:18-28
        self.critic_opt.zero_grad()
        values = self._values(observations)
        loss = F.mse_loss(values, returns)
        loss.backward()
        if self.gradient_clip > 0:                    # critics.py:24-25
            torch.nn.utils.clip_grad_norm_(list(self.critic.parameters()), self.gradient_clip)
        self.critic_opt.step()
        return dict(loss=loss.detach(), v=values.detach())

    def _actor_step(self, observations, actions, advantages, log_probs):
        ppo = self.kind == 'PPO'
        if (advantages == 0.).all():                  # actors.py:22-28 / 71-78
            zero = torch.as_tensor(0., dtype=torch.float32)
            with torch.no_grad():
                dist = gaussian_actor_forward(self.actor, observations)
                entropy, std = dist.entropy().mean(), dist.stddev.mean()
            out = dict(loss=zero, kl=zero, entropy=entropy)
            if ppo:
                out['clip_fraction'] = zero
            out['std'] = std
            if ppo:
                out['stop'] = zero > self.kl_threshold
            return out
        self.actor_opt.zero_grad()
        dist = gaussian_actor_forward(self.actor, observations)
        new_log_probs = dist.log_prob(actions).sum(dim=-1)
        if ppo:                                       # actors.py:84-90
            ratios = torch.exp(new_log_probs - log_probs)
            lo, hi = 1 - self.ratio_clip, 1 + self.ratio_clip
            clipped_ratios = torch.clamp(ratios, lo, hi)
            loss = -(torch.min(advantages * ratios, advantages * clipped_ratios)).mean()
        else:                                         # actors.py:34
            loss = -(advantages * new_log_probs).mean()
        entropy = dist.entropy().mean()
        if self.entropy_coeff != 0:
            loss -= self.entropy_coeff * entropy
        loss.backward()
        if self.gradient_clip > 0:                    # actors.py:37-38,96-98
            torch.nn.utils.clip_grad_norm_(list(self.actor.parameters()), self.gradient_clip)
        self.actor_opt.step()
        with torch.no_grad():
            kl = (log_probs - new_log_probs).mean()
        out = dict(loss=loss.detach(), kl=kl.detach(), entropy=entropy.detach())
        if ppo:                                       # actors.py:105-107
            out['clip_fraction'] = torch.as_tensor(
                ratios.gt(hi) | ratios.lt(lo), dtype=torch.float32).mean()
        out['std'] = dist.stddev.mean().detach()
        if ppo:
            out['stop'] = kl > self.kl_threshold
        return out

    def _log(self, group, infos):
        for k, v in infos.items():
            self.log(group + '/' + k, v.numpy())

    def _update(self):
        self._evaluate()
        if self.kind == 'PPO':                        # ppo.py:27-54
            train_actor, n_actor, n_critic = True, 0, 0
            keys = ('observations', 'actions', 'advantages', 'log_probs', 'returns')
            for batch in self.replay.batches(*keys):
                b = {k: torch.as_tensor(v) for k, v in batch.items()}
                infos = {}
                if train_actor:
                    infos['actor'] = self._actor_step(
                        b['observations'], b['actions'], b['advantages'], b['log_probs'])
                    n_actor += 1
                infos['critic'] = self._critic_step(b['observations'], b['returns'])
                n_critic += 1
                if train_actor:
                    train_actor = not infos['actor']['stop'].numpy()
                for group in infos:
                    self._log(group, infos[group])
            self.log('actor/iterations', n_actor)
            self.log('critic/iterations', n_critic)
        else:                                         # a2c.py:107-121
            b = self.replay.flat('observations', 'actions', 'advantages', 'log_probs')
            b = {k: torch.as_tensor(v) for k, v in b.items()}
            self._log('actor', self._actor_step(**b))
            for batch in self.replay.batches('observations', 'returns'):
                b = {k: torch.as_tensor(v) for k, v in batch.items()}
                self._log('critic', self._critic_step(**b))
        self.norm.update()                            # ppo.py:56-57, a2c.py:124-125


class OffPolicyOracle:
    """DDPG (torch/agents/ddpg.py:20-112), TD3 (td3.py:20-55), SAC (sac.py:22-51)."""

    def __init__(self, kind, hidden, buffer, start_steps=20000, log=None,
                 target_coeff=0.005, noise_scale=0.1, delay_steps=2,
                 entropy_coeff=0.2, target_noise=(0.2, 0.5), exploration='normal',
                 gradient_clip=0):
        self.gradient_clip = gradient_clip            # actors.py:176-177,256-257; critics.py:82-83,177-178,230-231
        self.kind, self.hidden = kind, tuple(hidden)
        self.exploration, self.ou_noises = exploration, None   # 'ou': noisy.py:53-88
        self.replay = RingStore(**buffer)
        self.start_steps, self.noise_scale = start_steps, noise_scale
        self.log = log or (lambda *a, **k: None)
        self.tau, self.delay_steps, self.alpha = target_coeff, delay_steps, entropy_coeff
        self.target_noise = target_noise
        self.lr = 3e-4 if kind == 'SAC' else 1e-3     # actors.py:162,229; critics.py:60,144,191

    def initialize(self, observation_space, action_space, seed=None):
        if seed is not None:
            np.random.seed(seed)
            import random
            random.seed(seed)
            torch.manual_seed(seed)
        obs, act = observation_space.shape[0], action_space.shape[0]
        self.act = act
        self.norm = RunningMoments(obs)
        make_actor = squashed_actor if self.kind == 'SAC' else deterministic_actor
        twin = self.kind != 'DDPG'
        names = ['critic_1', 'critic_2'] if twin else ['critic']
        # Creation order = actor_critics.py:51-56 / 102-114 (targets are fresh
        # networks, consuming the RNG, then overwritten by assign_targets).
        self.actor = make_actor('actor.', obs, act, self.hidden)
        self.critics = [value_critic(n + '.', obs + act, self.hidden) for n in names]
        self.target_actor = make_actor('target_actor.', obs, act, self.hidden)
        self.target_critics = [value_critic('target_' + n + '.', obs + act, self.hidden)
                               for n in names]
        self.target_actor.freeze()
        self.target_actor.copy_from(self.actor)
        for t, o in zip(self.target_critics, self.critics):
            t.freeze()
            t.copy_from(o)
        self.replay.initialize(seed)
        self.noise_rng = np.random.RandomState(seed)   # explorations/noisy.py:13,36
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), lr=self.lr)
        critic_params = [p for c in self.critics for p in c.parameters()]
        self.critic_opt = torch.optim.Adam(critic_params, lr=self.lr)

    def state_dict(self):
        out = {}
        nets = [self.actor] + self.critics + [self.target_actor] + self.target_critics
        for net in nets:
            out.update(net.named())
            if 'critic' in net.prefix:
                out[net.prefix + 'encoder.observation_normalizer._mean'] = self.norm.t_mean
                out[net.prefix + 'encoder.observation_normalizer._std'] = self.norm.t_std
        out['observation_normalizer._mean'] = self.norm.t_mean
        out['observation_normalizer._std'] = self.norm.t_std
        if self.kind == 'TD3':     # td3.py:36 registers critic_1 again as `critic`
            for k in [k for k in out if k.startswith('critic_1.')]:
                out['critic.' + k[len('critic_1.'):]] = out[k]
        return out

    # -- acting -------------------------------------------------------------
    def _greedy(self, observations):                  # ddpg.py:78-81, sac.py:48-51
        with torch.no_grad():
            if self.kind == 'SAC':
                return torch.tanh(squashed_actor_forward(self.actor, as_f32(observations)).mean)
            return deterministic_actor_forward(self.actor, as_f32(observations))

    def _policy(self, observations):                  # ddpg.py:83-84, sac.py:40-46
        if self.kind == 'SAC':
            with torch.no_grad():
                dist = squashed_actor_forward(self.actor, as_f32(observations))
                return torch.tanh(dist.sample()).numpy()
        return self._greedy(observations).numpy()

    def step(self, observations, steps):              # ddpg.py:45-53 + noisy.py:15-22,38-47
        if steps > self.start_steps:
            actions = self._policy(observations)
            if self.exploration == 'ou':              # noisy.py:71-81
                scale, clip, theta, dt = 0.1, 2, .15, 1e-2          # defaults noisy.py:55
                if self.ou_noises is None:
                    self.ou_noises = np.zeros_like(actions)
                noises = self.noise_rng.normal(size=actions.shape)
                noises = np.clip(noises, -clip, clip)
                self.ou_noises -= theta * self.ou_noises * dt
                self.ou_noises += scale * np.sqrt(dt) * noises
                actions = (actions + self.ou_noises).astype(np.float32)
            elif self.kind != 'SAC':
                noises = self.noise_scale * self.noise_rng.normal(size=actions.shape)
                actions = (actions + noises).astype(np.float32)
            actions = np.clip(actions, -1, 1)
        else:
            actions = self.noise_rng.uniform(-1, 1, (len(observations), self.act))
        self.last_observations = np.array(observations, copy=True)
        self.last_actions = actions.copy()
        return actions

    def test_step(self, observations, steps):
        return self._greedy(observations).numpy()

    # -- learning -----------------------------------------------------------
    def update(self, observations, rewards, resets, terminations, steps):
        # ddpg.py:59-76
        self.replay.store(
            observations=self.last_observations, actions=self.last_actions,
            next_observations=observations, rewards=rewards, resets=resets,
            terminations=terminations)
        self.norm.record(self.last_observations)
        if self.replay.ready(steps):
            self._update(steps)
        if self.exploration == 'ou' and self.ou_noises is not None:   # ddpg.py:76 + noisy.py:86-88
            self.ou_noises *= (1. - resets)[:, None]

    def _q(self, net, observations, actions):
        return value_critic_forward(net, self.norm, torch.relu, observations, actions)

    def _critic_step(self, observations, actions, next_observations, rewards, discounts):
        with torch.no_grad():
            if self.kind == 'SAC':                    # critics.py:205-220
                dist = squashed_actor_forward(self.actor, next_observations)
                next_actions, next_logp = squashed_rsample_with_log_prob(dist)
                next_logp = next_logp.sum(dim=-1)
            else:                                     # critics.py:71-75,159-167
                next_actions = deterministic_actor_forward(self.target_actor, next_observations)
                if self.kind == 'TD3':                # critics.py:130-134
                    scale, clip = self.target_noise
                    noises = scale * torch.randn_like(next_actions)
                    noises = torch.clamp(noises, -clip, clip)
                    next_actions = torch.clamp(next_actions + noises, -1, 1)
            next_values = [self._q(t, next_observations, next_actions)
                           for t in self.target_critics]
            next_value = next_values[0] if len(next_values) == 1 else \
                torch.min(next_values[0], next_values[1])
            if self.kind == 'SAC':
                returns = rewards + discounts * (next_value - self.alpha * next_logp)
            else:
                returns = rewards + discounts * next_value
        self.critic_opt.zero_grad()
        values = [self._q(c, observations, actions) for c in self.critics]
        losses = [F.mse_loss(v, returns) for v in values]
        loss = losses[0] if len(losses) == 1 else losses[0] + losses[1]
        loss.backward()
        if self.gradient_clip > 0:                    # joint norm over all critics' variables
            torch.nn.utils.clip_grad_norm_(
                [p for c in self.critics for p in c.parameters()], self.gradient_clip)
        self.critic_opt.step()
        out = dict(loss=loss.detach())
        if len(values) == 1:
            out['q'] = values[0].detach()
        else:
            out['q1'], out['q2'] = values[0].detach(), values[1].detach()
        return out

    def _actor_step(self, observations):
        critic_params = [p for c in self.critics for p in c.parameters()]
        if self.kind == 'TD3':     # td3.py:36: model.critic = critic_1 only
            critic_params = self.critics[0].parameters()
        for p in critic_params:
            p.requires_grad = False
        self.actor_opt.zero_grad()
        if self.kind == 'SAC':                        # actors.py:238-267
            dist = squashed_actor_forward(self.actor, observations)
            actions, logp = squashed_rsample_with_log_prob(dist)
            logp = logp.sum(dim=-1)
            values = torch.min(self._q(self.critics[0], observations, actions),
                               self._q(self.critics[1], observations, actions))
            loss = (self.alpha * logp - values).mean()
        else:                                         # actors.py:170-189
            actions = deterministic_actor_forward(self.actor, observations)
            loss = -self._q(self.critics[0], observations, actions).mean()
        loss.backward()
        if self.gradient_clip > 0:
            torch.nn.utils.clip_grad_norm_(list(self.actor.parameters()), self.gradient_clip)
        self.actor_opt.step()
        for p in critic_params:
            p.requires_grad = True
        return dict(loss=loss.detach())

    def _soft_update(self):                           # actor_critics.py:68-72,126-130
        pairs = [(self.actor, self.target_actor)] + list(zip(self.critics, self.target_critics))
        with torch.no_grad():
            for online, target in pairs:
                for k in online.p:
                    target.p[k].data.mul_(1 - self.tau)
                    target.p[k].data.add_(self.tau * online.p[k].data)

    def _update(self, steps):                         # ddpg.py:86-112, td3.py:38-55
        keys = ('observations', 'actions', 'next_observations', 'rewards', 'discounts')
        for i, batch in enumerate(self.replay.batches(*keys, steps=steps)):
            b = {k: torch.as_tensor(v) for k, v in batch.items()}
            infos = dict(critic=self._critic_step(**b))
            if self.kind != 'TD3' or (i + 1) % self.delay_steps == 0:
                infos['actor'] = self._actor_step(b['observations'])
                self._soft_update()
            for group in infos:
                for k, v in infos[group].items():
                    self.log(group + '/' + k, v.numpy())
        self.norm.update()


def build(cfg, log=None):
    """Scenario config (oracle/scenarios.py) -> (agent, environment)."""
    env = VectorEnv(cfg['obs'], cfg['act'], cfg['workers'], cfg['max_episode_steps'],
                    time_feature=cfg.get('time_feature', False))
    env.initialize(cfg['seed'])
    if cfg['agent'] in ('PPO', 'A2C'):
        agent = OnPolicyOracle(cfg['agent'], cfg['hidden'], cfg['segment'], log=log,
                               gradient_clip=cfg.get('gradient_clip', 0))
    else:
        agent = OffPolicyOracle(cfg['agent'], cfg['hidden'], cfg['buffer'],
                                start_steps=cfg['start_steps'], log=log,
                                exploration=cfg.get('exploration', 'normal'),
                                gradient_clip=cfg.get('gradient_clip', 0))
    agent.initialize(env.observation_space, env.action_space, seed=cfg['seed'])
    return agent, env
