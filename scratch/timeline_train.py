"""Event timing of the fused train kernel vs forward + backward (critic minibatch, 16384 rows)."""
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
pool = torch.randn(rows * 32, 17, device='cuda')
idx = torch.randperm(rows * 32, device='cuda')[:rows]
targets = torch.randn(rows * 32, device='cuda')
stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.float64, device='cuda')
dout = torch.zeros(rows, 1, device='cuda')
out = torch.zeros(rows, 1, device='cuda')
inp = K.MlpInput(pool, idx=idx)
_lib.call('tb_tc_timeline', None)


def fused():
    net.train_step(inp, rows, dout, stats, idx=idx, targets=targets, out=out)


def chain():
    net.forward(inp, rows, out, save=True, vloss=(targets, idx, dout, stats))
    net.backward(dout, rows)


for name, fn in (('train kernel', fused), ('forward + backward', chain)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        fn()
    e.record()
    torch.cuda.synchronize()
    print(f'{name}: {s.elapsed_time(e) / 50 * 1e3:.1f} us per minibatch of {rows} rows')

fused()
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 64)()
_lib.call('tb_tc_timeline', buf)
t = np.array(buf[:], dtype=np.float64)
names = ['start', 'setup', 'x published', 'z1 complete', 'mid epilogue done', 'z2 complete', 'pass 1 done',
         'loss exchanged', 'pass 2 done', 'bwd GEMM complete', 'dz1 epilogue done']
print(' | '.join(f'{n} {(t[i] - t[0]) / 1965.0:.2f}' for i, n in enumerate(names)))
