"""Policy updaters (reference: tonic/torch/updaters/actors.py).

Each updater enqueues, without any host synchronisation, the kernel sequence
  forward (save activations) -> loss / head gradient -> backward -> weight
  gradients (split partial sums) -> Adam (+ partial reduction, transposes, flags)
for one minibatch addressed through an index vector.  Statistics are
accumulated in a float64 device block (layout: TB_STAT_* in the C header) and
read by the agent once per update.  `__call__` keeps the reference's
batch-tensor signature for direct use.
"""

import torch

from ... import _lib, config, kernels
from . import optimizers

FLOAT_EPSILON = 1e-8


def splits_for(rows):
    return max(1, min(config.wgrad_splits, rows // 128))


class _GaussianPolicyUpdater:
    ratio_clip = 0.0
    kl_threshold = -1.0

    def initialize(self, model):
        self.model = model
        actor = model.actor
        if actor.head.kind != 'detached_gaussian':
            raise NotImplementedError('policy-gradient kernels need the detached-scale '
                                      'Gaussian head (the reference default)')
        self.actor = actor
        self.variables = [p for p in actor.parameters() if p.requires_grad]
        self.adam = kernels.Adam(actor.network.params,
                                 **optimizers.adam_hyperparameters(self.optimizer, 3e-4))
        self.clipper = kernels.make_clipper(self.gradient_clip)
        self._rows = 0

    def _scratch(self, rows):
        if rows > self._rows:
            dev, A = kernels.device(), self.actor.action_size
            self._pre = torch.empty(rows, A, dtype=torch.float32, device=dev)
            self._dout = torch.empty(rows, 2 * A, dtype=torch.float32, device=dev)
            self._rows = rows
        return self._pre, self._dout

    def launch(self, observations, actions, advantages, log_probs, idx, rows, stats, stop=None,
               rows_global=None):
        """`rows` = transitions of this minibatch owned by this rank, `rows_global` =
        size of the whole minibatch (the mean's denominator)."""
        actor, net = self.actor, self.actor.network
        A = actor.action_size
        dout = None
        if rows > 0 and net.mlp.fused_train():
            # forward -> policy loss -> backward in one launch
            pre, dout = self._scratch(rows)
            net.mlp.train_step(actor.input(observations, idx), rows, dout, stats, idx=idx,
                               policy=dict(log_scale=net.extra('log_scale'), actions=actions,
                                           advantages=advantages, log_probs=log_probs,
                                           ratio_clip=self.ratio_clip, entropy_coeff=self.entropy_coeff),
                               out=pre, skip=stop)
        elif rows > 0:
            pre, dout = self._scratch(rows)
            actor.pre_activations(observations, out=pre, idx=idx, rows=rows, save=True, skip=stop)
            kernels.gauss_policy_loss(pre, net.extra('log_scale'), actions, advantages,
                                      log_probs, idx, rows, dout, stats, self.ratio_clip,
                                      self.entropy_coeff, skip=stop)
            net.mlp.backward(dout, rows, skip=stop)
        kernels.wgrad_and_apply(self.adam, net.mlp, dout, rows, rows_global or rows, n_extra=A,
                                off_extra=net.extra_offset('log_scale'), skip=stop, stats=stats,
                                kl_threshold=self.kl_threshold, stop=stop, clip=self.clipper)

    def infos(self, s):
        """Statistics block (host copy) -> the reference's info dict (python floats)."""
        rows = s[_lib.STAT_ROWS]
        A = self.actor.action_size
        trained = s[_lib.STAT_NONZERO_ADV] > 0
        out = dict(
            loss=(s[_lib.STAT_LOSS] / rows
                  - self.entropy_coeff * s[_lib.STAT_ENTROPY] / (rows * A)) if trained else 0.0,
            kl=s[_lib.STAT_KL] / rows if trained else 0.0,
            entropy=s[_lib.STAT_ENTROPY] / (rows * A))
        if self.ratio_clip > 0:
            out['clip_fraction'] = s[_lib.STAT_CLIPPED] / rows if trained else 0.0
        out['std'] = s[_lib.STAT_STD] / (rows * A)
        if self.ratio_clip > 0:
            out['stop'] = bool(out['kl'] > self.kl_threshold)
        return out

    def __call__(self, observations, actions, advantages, log_probs):
        args = [kernels.to_device(a) for a in (observations, actions, advantages, log_probs)]
        stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.float64, device=kernels.device())
        self.launch(*args, None, args[0].shape[0], stats)
        return {k: torch.as_tensor(v) for k, v in self.infos(stats.cpu().numpy()).items()}


class StochasticPolicyGradient(_GaussianPolicyUpdater):
    """A2C policy gradient, loss = -mean(advantages * log_prob)
    (reference: updaters/actors.py:9-50)."""

    def __init__(self, optimizer=None, entropy_coeff=0, gradient_clip=0):
        self.optimizer, self.entropy_coeff = optimizer, entropy_coeff
        self.gradient_clip = gradient_clip


class ClippedRatio(_GaussianPolicyUpdater):
    """PPO clipped surrogate with KL early stopping (reference:
    updaters/actors.py:53-112; the stop decision is taken on the device by the
    Adam kernel so the update loop never waits for the host)."""

    def __init__(self, optimizer=None, ratio_clip=0.2, kl_threshold=0.015, entropy_coeff=0,
                 gradient_clip=0):
        self.optimizer = optimizer
        self.ratio_clip, self.kl_threshold = ratio_clip, kl_threshold
        self.entropy_coeff = entropy_coeff
        self.gradient_clip = gradient_clip


class _CriticGradientUpdater:
    """Actor updates that differentiate through frozen critics (the reference sets
    `requires_grad = False` on the critic variables, actors.py:171-174,239-244;
    here the critics' weight-gradient and Adam kernels are simply not launched)."""

    default_lr = 1e-3

    def initialize(self, model):
        self.model = model
        self.actor = model.actor
        self.variables = [p for p in self.actor.parameters() if p.requires_grad]
        self.adam = kernels.Adam(self.actor.network.params,
                                 **optimizers.adam_hyperparameters(self.optimizer, self.default_lr))
        self.clipper = kernels.make_clipper(self.gradient_clip)
        self.obs_size = self.actor.network.layout.d_in
        self.seed, self._counter, self._rows = 0, 0, 0

    def _scratch(self, rows):
        if rows > self._rows:
            dev, A = kernels.device(), self.actor.action_size
            new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)   # noqa: E731
            self._pre = new(rows, self.actor.network.layout.n_out)
            self._actions, self._eps, self._logp = new(rows, A), new(rows, A), new(rows)
            self._q = [new(rows, 1), new(rows, 1)]
            self._dq = [new(rows, 1), new(rows, 1)]
            self._dqda = [new(rows, A), new(rows, A)]
            self._dout = new(rows, self.actor.network.layout.n_out)
            self._rows = rows

    def _finish(self, rows, stats, rows_global=None):
        net = self.actor.network
        if rows > 0:
            net.mlp.backward(self._dout, rows)
        kernels.wgrad_and_apply(self.adam, net.mlp, self._dout if rows > 0 else None, rows,
                                rows_global or rows, reduce_stats=stats, clip=self.clipper)

    @staticmethod
    def infos(s):
        return dict(loss=s[_lib.STAT_LOSS] / s[_lib.STAT_ROWS])


class DeterministicPolicyGradient(_CriticGradientUpdater):
    """loss = -mean Q(s, mu(s)) (reference: updaters/actors.py:159-189)."""

    def __init__(self, optimizer=None, gradient_clip=0):
        self.optimizer, self.gradient_clip = optimizer, gradient_clip

    def launch(self, observations, idx, rows, stats, rows_global=None, mine=None):
        if rows == 0:
            return self._finish(0, stats, rows_global)
        self._scratch(rows)
        actor, critic = self.actor, self.model.critic
        A = actor.action_size
        actor.pre_activations(observations, out=self._pre[:rows], idx=idx, rows=rows, save=True)
        kernels.tanh_action(self._pre[:rows], self._actions[:rows], mode=0)
        critic.values(observations, self._actions[:rows], out=self._q[0][:rows], idx=idx,
                      rows=rows, gather_actions=False, save=True)
        kernels.q_actor_loss(self._q[0], None, None, 0.0, rows, self._dq[0], None, stats)
        critic.network.mlp.backward(self._dq[0], rows, dx=self._dqda[0][:rows],
                                    dx_col0=self.obs_size)
        kernels.dpg_head_grad(self._dqda[0][:rows], self._actions[:rows], self._dout[:rows])
        self._finish(rows, stats, rows_global)

    def __call__(self, observations):
        observations = kernels.to_device(observations)
        stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.float64, device=kernels.device())
        self.launch(observations, None, observations.shape[0], stats)
        return {k: torch.as_tensor(v) for k, v in self.infos(kernels.to_host(stats)).items()}


class TwinCriticSoftDeterministicPolicyGradient(_CriticGradientUpdater):
    """SAC actor, loss = mean(alpha * log pi(a|s) - min(Q1, Q2)(s, a)), a ~ pi
    reparameterised (reference: updaters/actors.py:226-267)."""

    default_lr = 3e-4

    def __init__(self, optimizer=None, entropy_coeff=0.2, gradient_clip=0):
        self.optimizer, self.entropy_coeff = optimizer, entropy_coeff
        self.gradient_clip = gradient_clip

    def launch(self, observations, idx, rows, stats, rows_global=None, mine=None):
        A = self.actor.action_size
        eps = None
        if config.noise == 'host':      # rsample() of the whole batch, this rank's rows
            eps = torch.randn(rows_global or rows, A)
            eps = (eps if mine is None else eps[mine]).to(kernels.device())
        if rows == 0:
            return self._finish(0, stats, rows_global)
        self._scratch(rows)
        actor = self.actor
        critics = [self.model.critic_1, self.model.critic_2]
        actor.pre_activations(observations, out=self._pre[:rows], idx=idx, rows=rows, save=True)
        kernels.squashed_sample(self._pre[:rows], self._actions[:rows], self._logp[:rows],
                                eps=eps, eps_out=self._eps[:rows], seed=self.seed ^ 0xac7,
                                counter=self._counter)
        self._counter += rows
        for k, critic in enumerate(critics):
            critic.values(observations, self._actions[:rows], out=self._q[k][:rows], idx=idx,
                          rows=rows, gather_actions=False, save=True)
        kernels.q_actor_loss(self._q[0], self._q[1], self._logp, self.entropy_coeff, rows,
                             self._dq[0], self._dq[1], stats)
        for k, critic in enumerate(critics):
            critic.network.mlp.backward(self._dq[k], rows, dx=self._dqda[k][:rows],
                                        dx_col0=self.obs_size)
        kernels.sac_head_grad(self._pre[:rows], self._eps[:rows], self._actions[:rows],
                              self._dqda[0][:rows], self._dqda[1][:rows], self.entropy_coeff,
                              self._dout[:rows])
        self._finish(rows, stats, rows_global)
