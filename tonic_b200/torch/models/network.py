"""Binding between the reference-shaped module tree and the flat device buffers
the kernels work on."""

import torch

from ... import kernels


class BoundNetwork:
    """A torso MLP + head living in one `kernels.DeviceMlp`."""

    def __init__(self, torso, head_linears, extras=()):
        linears = torso.linears()
        if len(linears) != 2 or linears[0].out_features != linears[1].out_features:
            raise NotImplementedError(
                'the sm_100a MLP kernels implement two hidden layers of equal width '
                f'(got sizes {torso.sizes})')
        hidden = linears[0].out_features
        if hidden not in (64, 128, 256):
            raise NotImplementedError(f'hidden width {hidden} not in (64, 128, 256)')
        d_in = linears[0].in_features
        n_out = sum(l.out_features for l in head_linears)
        layout = kernels.MlpLayout(d_in, hidden, n_out, torso.activation_name(),
                                   [(name, p.numel()) for name, p in extras])
        self.mlp = kernels.DeviceMlp(layout)
        self.layout = layout
        m = self.mlp
        # copy the CPU-initialised values into the flat buffer, then re-home the
        # parameters as views of it (kernels update them in place)
        self._rehome(linears[0].weight, m.view('w1', (hidden, d_in)))
        self._rehome(linears[0].bias, m.view('b1', (hidden,)))
        self._rehome(linears[1].weight, m.view('w2', (hidden, hidden)))
        self._rehome(linears[1].bias, m.view('b2', (hidden,)))
        w3, b3 = m.view('w3', (n_out, hidden)), m.view('b3', (n_out,))
        row = 0
        for l in head_linears:
            self._rehome(l.weight, w3[row:row + l.out_features])
            self._rehome(l.bias, b3[row:row + l.out_features])
            row += l.out_features
        for name, p in extras:
            self._rehome(p, m.view(name, tuple(p.shape)))
        m.pack()

    @staticmethod
    def _rehome(param, view):
        view.copy_(param.data)
        param.data = view

    @property
    def params(self):
        return self.mlp.params

    def extra_offset(self, name):
        return self.layout.offsets[name][0]

    def extra(self, name):
        off, size = self.layout.offsets[name]
        return self.mlp.params[off:off + size]

    def refresh(self):
        """Call after parameters were written from outside (checkpoint load)."""
        self.mlp.pack()

    def copy_from(self, other):
        self.mlp.params.copy_(other.mlp.params)
        self.mlp.packed.copy_(other.mlp.packed)

    def soft_update_from(self, other, tau):
        kernels.soft_update(self.mlp.params, other.mlp.params, tau)
        # transposes / tf32 splits are rebuilt from the updated parameters (a blend of
        # two tf32-exact numbers is not tf32-exact, so the splits cannot be blended)
        self.mlp.pack()


def scratch(rows, cols, like):
    return torch.empty(rows, cols, dtype=torch.float32, device=like.device)
