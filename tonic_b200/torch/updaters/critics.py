"""Value updaters (reference: tonic/torch/updaters/critics.py)."""

import torch

from ... import _lib, kernels
from . import optimizers
from .actors import splits_for


class VRegression:
    """MSE regression of V(s) on the lambda-returns (reference:
    updaters/critics.py:6-28)."""

    def __init__(self, loss=None, optimizer=None, gradient_clip=0):
        if loss is not None and not isinstance(loss, torch.nn.MSELoss):
            raise NotImplementedError('only the MSE loss has a kernel')
        self.optimizer, self.gradient_clip = optimizer, gradient_clip

    def initialize(self, model):
        self.model = model
        self.clipper = kernels.make_clipper(self.gradient_clip)
        self.critic = model.critic
        self.variables = [p for p in self.critic.parameters() if p.requires_grad]
        self.adam = kernels.Adam(self.critic.network.params,
                                 **optimizers.adam_hyperparameters(self.optimizer, 1e-3))
        self._rows = 0

    def _scratch(self, rows):
        if rows > self._rows:
            dev = kernels.device()
            self._values = torch.empty(rows, 1, dtype=torch.float32, device=dev)
            self._dout = torch.empty(rows, 1, dtype=torch.float32, device=dev)
            self._rows = rows
        return self._values, self._dout

    def launch(self, observations, returns, idx, rows, stats, rows_global=None):
        critic, net = self.critic, self.critic.network
        dout = None
        if rows > 0 and net.mlp.fused_train():
            # forward -> squared-error loss -> backward in one launch
            values, dout = self._scratch(rows)
            net.mlp.train_step(critic.input(observations, None, idx), rows, dout, stats, idx=idx,
                               targets=returns, out=values)
        elif rows > 0:
            values, dout = self._scratch(rows)
            # the tensor-core forward kernel evaluates the squared-error loss in its epilogue
            critic.values(observations, out=values, idx=idx, rows=rows, save=True,
                          vloss=(returns, idx, dout, stats))
            if not net.mlp.vloss_fused:
                kernels.mse_loss(values, returns, idx, rows, dout, stats)
            net.mlp.backward(dout, rows)
        kernels.wgrad_and_apply(self.adam, net.mlp, dout, rows, rows_global or rows,
                                reduce_stats=stats, clip=self.clipper)

    @staticmethod
    def infos(s):
        rows = s[_lib.STAT_ROWS]
        return dict(loss=s[_lib.STAT_LOSS] / rows, v=s[_lib.STAT_VALUE] / rows)

    def __call__(self, observations, returns):
        observations, returns = kernels.to_device(observations), kernels.to_device(returns)
        stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.float64, device=kernels.device())
        rows = observations.shape[0]
        self.launch(observations, returns, None, rows, stats)
        s = stats.cpu().numpy()
        return dict(loss=torch.as_tensor(s[_lib.STAT_LOSS] / rows),
                    v=self._values[:rows, 0].clone())


class _QLearning:
    """Shared launch sequence of the Q-learning updaters: target computation with
    the target networks, then one MSE step per critic on the same targets."""

    default_lr = 1e-3

    def __init__(self, loss=None, optimizer=None, gradient_clip=0):
        if loss is not None and not isinstance(loss, torch.nn.MSELoss):
            raise NotImplementedError('only the MSE loss has a kernel')
        self.optimizer, self.gradient_clip = optimizer, gradient_clip

    def _critics(self, model):
        raise NotImplementedError

    def initialize(self, model):
        self.model = model
        self.critics, self.target_critics = self._critics(model)
        self.variables = [p for c in self.critics for p in c.parameters() if p.requires_grad]
        hyper = optimizers.adam_hyperparameters(self.optimizer, self.default_lr)
        # one torch Adam over both critics == one Adam per critic with equal step counts
        self.adams = [kernels.Adam(c.network.params, **hyper) for c in self.critics]
        # one optimizer over the variables of all critics: the clip norm is their JOINT norm
        self.clipper = kernels.make_clipper(self.gradient_clip)
        if self.clipper:
            self.clipper.deferred = True
        self.action_size = model.actor.action_size
        self.seed, self._counter, self._rows = 0, 0, 0

    def _scratch(self, rows):
        if rows > self._rows:
            dev, A = kernels.device(), self.action_size
            new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)   # noqa: E731
            self._next_pre = new(rows, self.model.actor.network.layout.n_out)
            self._next_actions = new(rows, A)
            self._next_logp = new(rows)
            self._next_q = [new(rows, 1), new(rows, 1)]
            self._targets = new(rows)
            self._values = [new(rows, 1), new(rows, 1)]
            self._dout = new(rows, 1)
            self._rows = rows

    def next_actions(self, next_observations, idx, rows, rows_global, mine):
        """Fills self._next_actions; returns the log-probs to subtract or None."""
        raise NotImplementedError

    def skip_noise(self, rows_global):
        """Keeps the host noise stream aligned when this rank owns no sample."""

    def host_noise(self, rows, rows_global, mine, device):
        """[rows, A] standard normals from torch's global CPU generator: the block
        the single-process reference would draw for the whole batch, restricted to
        the samples this rank owns."""
        noise = torch.randn(rows_global, self.action_size)
        if mine is not None:
            noise = noise[mine]
        return noise.to(device)

    entropy_coeff = 0.0

    def launch(self, replay, idx, rows, stats, rows_global=None, mine=None):
        if rows == 0:
            self.skip_noise(rows_global)
            for k, critic in enumerate(self.critics):
                kernels.apply_gradients(self.adams[k], critic.network.mlp, None, 1, 0,
                                        rows_global,
                                        reduce_stats=stats if k == len(self.critics) - 1 else None,
                                        clip=self.clipper)
            if self.clipper:
                self.clipper.finish()
            return
        self._scratch(rows)
        obs, acts = replay.flat('observations'), replay.flat('actions')
        nobs = replay.flat('next_observations')
        next_logp = self.next_actions(nobs, idx, rows, rows_global or rows, mine)
        for k, target in enumerate(self.target_critics):        # critics.py:73-74,162-166,214-218
            target.values(nobs, self._next_actions[:rows], out=self._next_q[k][:rows], idx=idx,
                          rows=rows, gather_actions=False)
        q2 = self._next_q[1] if len(self.target_critics) > 1 else None
        kernels.q_target(replay.flat('rewards'), replay.flat('terminations'), idx,
                         replay.discount_factor, self._next_q[0], q2, next_logp,
                         self.entropy_coeff, rows, self._targets,
                         discounts=replay.flat('discounts') if replay.return_steps > 1 else None)
        for k, critic in enumerate(self.critics):               # critics.py:77-84,169-179
            net = critic.network
            critic.values(obs, acts, out=self._values[k][:rows], idx=idx, rows=rows, save=True)
            kernels.mse_loss(self._values[k], self._targets, None, rows, self._dout, stats,
                             stat_slot=_lib.STAT_VALUE if k == 0 else _lib.STAT_VALUE2,
                             count_rows=(k == 0))
            net.mlp.backward(self._dout, rows)
            # the statistics block is all-reduced once (with the last critic)
            kernels.wgrad_and_apply(self.adams[k], net.mlp, self._dout, rows, rows_global or rows,
                                    reduce_stats=stats if k == len(self.critics) - 1 else None,
                                    clip=self.clipper)
        if self.clipper:
            self.clipper.finish()

    def infos(self, s):
        rows = s[_lib.STAT_ROWS]
        out = dict(loss=s[_lib.STAT_LOSS] / rows)
        if len(self.critics) == 1:
            out['q'] = s[_lib.STAT_VALUE] / rows
        else:
            out['q1'] = s[_lib.STAT_VALUE] / rows
            out['q2'] = s[_lib.STAT_VALUE2] / rows
        return out


class DeterministicQLearning(_QLearning):
    """DDPG critic (reference: updaters/critics.py:56-86)."""

    def _critics(self, model):
        return [model.critic], [model.target_critic]

    def next_actions(self, next_observations, idx, rows, rows_global, mine):
        pre = self._next_pre[:rows]
        self.model.target_actor.pre_activations(next_observations, out=pre, idx=idx, rows=rows)
        kernels.tanh_action(pre, self._next_actions[:rows], mode=0)
        return None


class TargetActionNoise:
    """Clipped Gaussian noise on the target actions (reference:
    updaters/critics.py:125-134); applied inside csrc/offpolicy.cu::tanh_action_kernel."""

    def __init__(self, scale=0.2, clip=0.5):
        self.scale, self.clip = scale, clip


class TwinCriticDeterministicQLearning(_QLearning):
    """TD3 critics (reference: updaters/critics.py:137-182)."""

    def __init__(self, loss=None, optimizer=None, target_action_noise=None, gradient_clip=0):
        super().__init__(loss, optimizer, gradient_clip)
        self.target_action_noise = target_action_noise or TargetActionNoise(scale=0.2, clip=0.5)

    def _critics(self, model):
        return [model.critic_1, model.critic_2], [model.target_critic_1, model.target_critic_2]

    def skip_noise(self, rows_global):
        from ... import config
        if config.noise == 'host':
            torch.randn(rows_global, self.action_size)

    def next_actions(self, next_observations, idx, rows, rows_global, mine):
        from ... import config
        A = self.action_size
        pre = self._next_pre[:rows]
        self.model.target_actor.pre_activations(next_observations, out=pre, idx=idx, rows=rows)
        noise = None
        if config.noise == 'host':      # torch.randn_like(next_actions), critics.py:131
            noise = self.host_noise(rows, rows_global, mine, pre.device)
        kernels.tanh_action(pre, self._next_actions[:rows], mode=1, noise32=noise,
                            seed=self.seed ^ 0x7d3, counter=self._counter,
                            noise_scale=self.target_action_noise.scale,
                            noise_clip=self.target_action_noise.clip)
        self._counter += rows
        return None


class TwinCriticSoftQLearning(_QLearning):
    """SAC critics (reference: updaters/critics.py:185-235): next actions from the
    ONLINE actor, soft target min(Q1', Q2') - alpha * log pi."""

    default_lr = 3e-4

    def __init__(self, loss=None, optimizer=None, entropy_coeff=0.2, gradient_clip=0):
        super().__init__(loss, optimizer, gradient_clip)
        self.entropy_coeff = entropy_coeff

    def _critics(self, model):
        return [model.critic_1, model.critic_2], [model.target_critic_1, model.target_critic_2]

    skip_noise = TwinCriticDeterministicQLearning.skip_noise

    def next_actions(self, next_observations, idx, rows, rows_global, mine):
        from ... import config
        A = self.action_size
        pre = self._next_pre[:rows]
        self.model.actor.pre_activations(next_observations, out=pre, idx=idx, rows=rows)
        eps = self.host_noise(rows, rows_global, mine, pre.device) \
            if config.noise == 'host' else None
        kernels.squashed_sample(pre, self._next_actions[:rows], self._next_logp[:rows], eps=eps,
                                seed=self.seed ^ 0x5ac, counter=self._counter)
        self._counter += rows
        return self._next_logp
