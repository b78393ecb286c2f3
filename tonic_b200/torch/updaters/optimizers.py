"""Translation of the reference's optimizer factories (`lambda params:
torch.optim.Adam(params, lr=...)`, updaters/actors.py:11-12 etc.) into the
hyper-parameters of the fused Adam kernel (csrc/optim.cu)."""

import torch


def adam_hyperparameters(factory, default_lr):
    if factory is None:
        return dict(lr=default_lr, betas=(0.9, 0.999), eps=1e-8)
    probe = factory([torch.nn.Parameter(torch.zeros(1))])
    if not isinstance(probe, torch.optim.Adam):
        raise NotImplementedError(f'only torch.optim.Adam has a kernel (got {type(probe)})')
    group = probe.param_groups[0]
    if group.get('weight_decay', 0) != 0 or group.get('amsgrad', False):
        raise NotImplementedError('Adam weight_decay / amsgrad are not implemented')
    return dict(lr=group['lr'], betas=tuple(group['betas']), eps=group['eps'])
