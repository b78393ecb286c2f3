"""Per-phase clock64() timeline of the fused forward kernel (CTA 0) + event timings."""
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
x = torch.randn(rows, 17, device='cuda')
inp = K.MlpInput(x)
out = torch.empty(rows, 1, device='cuda')
xin = torch.empty(rows, layout.ldx, device='cuda')
bufs = [torch.empty(rows, 256, device='cuda') for _ in range(3)]
_lib.call('tb_tc_timeline', None)


def run(save):
    b = bufs if save else [None] * 3
    _lib.call('tb_tc_mlp_forward', ctypes.byref(layout.shape), K.ptr(net.params), K.ptr(net.packed),
              ctypes.byref(inp.struct), rows, K.ptr(out), K.ptr(xin) if save else None, K.ptr(b[0]),
              K.ptr(b[1]), K.ptr(b[2]), 3, None, K.stream())


for save in (True, False):
    for _ in range(5):
        run(save)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        run(save)
    e.record()
    torch.cuda.synchronize()
    print(f'save={save}: {s.elapsed_time(e) / 50 * 1e3:.1f} us per launch (back to back)')
    buf = (ctypes.c_uint64 * 64)()
    _lib.call('tb_tc_timeline', buf)
    t = np.array(buf[:], dtype=np.float64)
    t0 = t[0]
    us = lambda i: (t[i] - t0) / 1965.0      # noqa: E731
    print(f'  setup done {us(1):.2f} | L1 operands written {us(2):.2f} | L1 acc complete {us(3):.2f}')
    for c in range(8):
        print(f'  chunk {c}: computed {us(4 + c):.2f}  stage free {us(12 + c):.2f}  published {us(20 + c):.2f}'
              f' | MMA: B landed {us(41 + c):.2f}  issue {us(51 + c):.2f}')
    print(f'  MMA L1: B {us(40):.2f} issue {us(50):.2f}')
    print(f'  mid done {us(28):.2f} | L2 acc complete {us(29):.2f} | final epilogue done {us(30):.2f}')
