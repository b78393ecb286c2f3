"""clock64() timeline of the fused backward kernel (CTA 0)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tonic_b200 import _lib, kernels as K  # noqa: E402

K.device()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
layout = K.MlpLayout(17, 256, 1, 'tanh')
net = K.DeviceMlp(layout)
net.params.copy_(torch.randn(layout.n_params) * 0.15)
net.pack()
h1 = torch.tanh(torch.randn(rows, 256, device='cuda'))
h1_hi, h1_lo = torch.empty_like(h1), torch.empty_like(h1)
K.split_tf32(h1, h1_hi, h1_lo)
h2 = torch.tanh(torch.randn(rows, 256, device='cuda'))
dout = torch.randn(rows, 1, device='cuda')
outs = [torch.empty(rows, 256, device='cuda') for _ in range(3)]
_lib.call('tb_tc_timeline', None)


def run():
    _lib.call('tb_tc_mlp_backward', ctypes.byref(layout.shape), K.ptr(net.params), K.ptr(net.packed),
              K.ptr(dout), 1, K.ptr(h1_hi), K.ptr(h1_lo), K.ptr(h2), rows, K.ptr(outs[0]),
              K.ptr(outs[1]), K.ptr(outs[2]), 3, None, K.stream())


for _ in range(5):
    run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(50):
    run()
e.record()
torch.cuda.synchronize()
print(f'{s.elapsed_time(e) / 50 * 1e3:.1f} us per launch (back to back)')
buf = (ctypes.c_uint64 * 64)()
_lib.call('tb_tc_timeline', buf)
t = np.array(buf[:], dtype=np.float64)
us = lambda i: (t[i] - t[0]) / 1965.0      # noqa: E731
print(f'setup done {us(1):.2f}')
for c in range(8):
    print(f'chunk {c}: computed {us(4 + c):.2f}  stage free {us(12 + c):.2f}  published {us(20 + c):.2f}'
          f' | MMA: B landed {us(40 + c):.2f}  issue {us(50 + c):.2f}')
for wg in range(2):
    print(f'group {wg}: dz2 phase done {us(28 + 4 * wg):.2f} | acc complete {us(29 + 4 * wg):.2f} | '
          f'epilogue done {us(30 + 4 * wg):.2f}')
