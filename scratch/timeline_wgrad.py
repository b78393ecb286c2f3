"""clock64() timeline of the fused weight-gradient kernel (CTA 0,0) + event timing."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tonic_b200 import _lib, kernels as K  # noqa: E402

K.device()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n_out, n_extra = (6, 6) if 'actor' in sys.argv else (1, 0)
PLAIN = 'plain' in sys.argv
PASSES = 1 if 'p1' in sys.argv else 3
extras = [('log_scale', n_extra)] if n_extra else ()
layout = K.MlpLayout(17, 256, n_out, 'tanh', extras)
net = K.DeviceMlp(layout)
net.params.copy_(torch.randn(layout.n_params) * 0.1)
net.pack()
adam = K.Adam(net.params, lr=1e-3)
dev = lambda *s: torch.randn(*s, device='cuda') * 0.1   # noqa: E731
xin, h2, dz1, dout = dev(rows, layout.ldx), dev(rows, 256), dev(rows, 256), dev(rows, n_out + n_extra)
h1_hi, h1_lo, dz2_hi, dz2_lo = dev(rows, 256), dev(rows, 256) * 1e-3, dev(rows, 256), dev(rows, 256) * 1e-3
n_split = 74
gpart = torch.zeros(n_split, layout.n_params, device='cuda')
flat = torch.zeros(layout.n_params, device='cuda')
sync = torch.zeros(1, dtype=torch.int64, device='cuda')
off_extra = layout.offsets['log_scale'][0] if n_extra else 0
_lib.call('tb_wgrad_timeline', None)


def run(fuse):
    _lib.call('tb_mlp_wgrad_fused', ctypes.byref(layout.shape), K.ptr(xin), K.ptr(h1_hi), None if PLAIN else K.ptr(h1_lo),
              K.ptr(h2), K.ptr(dz1), K.ptr(dz2_hi), None if PLAIN else K.ptr(dz2_lo), K.ptr(dout), n_out + n_extra, n_extra,
              off_extra, rows, K.ptr(gpart), n_split, K.ptr(flat), K.ptr(sync), PASSES,
              ctypes.byref(adam.struct) if fuse else None, K.ptr(net.packed) if fuse else None,
              1.0 / rows, None, -1.0, None, None, None, None, None, K.stream())


for fuse in (False, True):
    for _ in range(5):
        run(fuse)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        run(fuse)
    e.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * 64)()
    _lib.call('tb_wgrad_timeline', buf)
    t = np.array(buf[:], dtype=np.float64)
    us = lambda i: (t[i] - t[0]) / 1965.0      # noqa: E731
    print(f'passes={PASSES} fuse_adam={fuse} plain={PLAIN} rows={rows} n_out={n_out}: {s.elapsed_time(e) / 50 * 1e3:.1f} us per launch')
    print(f'  (from setup done) MMAs issued {us(1):.2f} | accumulator complete {us(2):.2f} | narrow done {us(3):.2f}'
          f' | partials written {us(4):.2f} | at barrier {us(5):.2f} | barrier passed {us(6):.2f} | reduced {us(7):.2f}')
    print('  chunk: TMA issue / landed / MMA start: ' + '  '.join(f'{c}: {us(16 + c):.2f}/{us(32 + c):.2f}/{us(48 + c):.2f}' for c in range(14)))
