"""MLP torso and parameter helpers (reference: tonic/torch/models/utils.py:4-27).

Layers are created as ordinary `torch.nn.Linear` modules on the CPU -- in the
reference's creation order, so the default initialisation consumes torch's global
generator exactly like the reference -- and are then re-homed into one flat
float32 CUDA buffer per network by `network.bind` (models/network.py).  The
modules keep the reference's names, so `state_dict()` has the same keys and a
reference checkpoint loads unchanged.
"""

import torch


class MLP(torch.nn.Module):
    def __init__(self, sizes, activation, fn=None):
        super().__init__()
        self.sizes = tuple(sizes)
        self.activation = activation
        self.fn = fn

    def initialize(self, input_size):
        widths = [input_size] + list(self.sizes)
        layers = []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            layers.append(torch.nn.Linear(fan_in, fan_out))
            layers.append(self.activation())
        self.model = torch.nn.Sequential(*layers)
        if self.fn is not None:
            self.model.apply(self.fn)
        return widths[-1]

    def linears(self):
        return [m for m in self.model if isinstance(m, torch.nn.Linear)]

    def activation_name(self):
        if self.activation is torch.nn.Tanh:
            return 'tanh'
        if self.activation is torch.nn.ReLU:
            return 'relu'
        raise NotImplementedError(
            f'activation {self.activation} has no sm_100a kernel (Tanh and ReLU do)')

    def forward(self, inputs):
        raise RuntimeError('device MLPs run through their owning Actor / Critic')


def trainable_variables(model):
    return [p for p in model.parameters() if p.requires_grad]
