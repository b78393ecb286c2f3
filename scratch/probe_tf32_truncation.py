"""Does tcgen05 kind::tf32 truncate its float32 operands?  Fused forward with the hi tiles holding
masked (tf32-exact) values vs plain float32 values: bit-identical outputs <=> truncation."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from tonic_b200 import _lib, kernels as K  # noqa: E402

K.device()
rows = 4096
layout = K.MlpLayout(17, 256, 6, 'tanh')
net = K.DeviceMlp(layout)
net.params.copy_(torch.randn(layout.n_params) * 0.2)
net.pack()
x = torch.randn(rows, 17, device='cuda')
inp = K.MlpInput(x)
outs = []
for mode in (0, 1):
    _lib.call('tb_debug_plain_hi', mode)
    out = torch.empty(rows, 6, device='cuda')
    bufs = [torch.empty(rows, 256, device='cuda') for _ in range(3)]
    xin = torch.empty(rows, layout.ldx, device='cuda')
    _lib.call('tb_tc_mlp_forward', ctypes.byref(layout.shape), K.ptr(net.params), K.ptr(net.packed),
              ctypes.byref(inp.struct), rows, K.ptr(out), K.ptr(xin), K.ptr(bufs[0]), K.ptr(bufs[1]),
              K.ptr(bufs[2]), 3, None, K.stream())
    torch.cuda.synchronize()
    outs.append((out.clone(), bufs[2].clone(), bufs[0].clone()))
_lib.call('tb_debug_plain_hi', 0)
print('head outputs bit-identical:', torch.equal(outs[0][0], outs[1][0]))
print('h2 bit-identical:', torch.equal(outs[0][1], outs[1][1]))
print('max |diff| out', (outs[0][0] - outs[1][0]).abs().max().item(), 'h2', (outs[0][1] - outs[1][1]).abs().max().item())
print('saved hi tile differs (plain vs masked), as expected:', not torch.equal(outs[0][2], outs[1][2]))
