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
        if gradient_clip:
            raise NotImplementedError('gradient clipping is not implemented')
        self.optimizer = optimizer

    def initialize(self, model):
        self.model = model
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

    def launch(self, observations, returns, idx, rows, stats):
        critic, net = self.critic, self.critic.network
        values, dout = self._scratch(rows)
        n_split = splits_for(rows)
        critic.values(observations, out=values, idx=idx, rows=rows, save=True)
        kernels.mse_loss(values, returns, idx, rows, dout, stats)
        net.mlp.backward(dout, rows)
        gpart = net.mlp.wgrad(dout, rows, n_split)
        self.adam.step(net.mlp, gpart, n_split, 1.0 / rows)

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
