"""torchrun: per-launch time and CTA (0,0) timeline of the fused weight-gradient kernel with the
gradient exchange between ranks inside (vs. the same launch without peers)."""
import ctypes
import faulthandler
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
faulthandler.dump_traceback_later(150, exit=True)
local_rank = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
from tonic_b200 import _lib, distributed, kernels as K  # noqa: E402

K.device()
rows = 16384
rank = dist.get_rank()
for n_out, n_extra in ((1, 0), (6, 6)):
    extras = [('log_scale', n_extra)] if n_extra else ()
    layout = K.MlpLayout(17, 256, n_out, 'tanh', extras)
    net = K.DeviceMlp(layout)
    net.params.copy_(torch.randn(layout.n_params) * 0.1)
    net.pack()
    adam = K.Adam(net.params, lr=1e-3)
    dev = lambda *s: torch.randn(*s, device='cuda') * 0.1   # noqa: E731
    xin, h2, dz1, dout = dev(rows, layout.ldx), dev(rows, 256), dev(rows, 256), dev(rows, n_out + n_extra)
    h1, dz2 = dev(rows, 256), dev(rows, 256)
    n_split = 74
    gpart = torch.zeros(n_split, layout.n_params, device='cuda')
    flat = torch.zeros(layout.n_params, device='cuda')
    sync = torch.zeros(1, dtype=torch.int64, device='cuda')
    stats = torch.ones(12, dtype=torch.float64, device='cuda')
    off_extra = layout.offsets['log_scale'][0] if n_extra else 0
    region = distributed.PeerRegion(layout.n_params, fused=True)
    _lib.call('tb_wgrad_timeline', None)

    def run(peers):
        _lib.call('tb_mlp_wgrad_fused', ctypes.byref(layout.shape), K.ptr(xin), K.ptr(h1), None,
                  K.ptr(h2), K.ptr(dz1), K.ptr(dz2), None, K.ptr(dout), n_out + n_extra, n_extra,
                  off_extra, rows, K.ptr(gpart), n_split, K.ptr(flat), K.ptr(sync), 3,
                  ctypes.byref(adam.struct), K.ptr(net.packed), 1.0 / rows, None, -1.0, None, None,
                  ctypes.byref(region.struct) if peers else None, K.ptr(region.epoch) if peers else None,
                  K.ptr(stats) if peers else None, K.stream())

    for peers in (False, True, False, True):
        for _ in range(10):
            run(peers)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(200):
            run(peers)
        e.record()
        torch.cuda.synchronize()
        buf = (ctypes.c_uint64 * 64)()
        _lib.call('tb_wgrad_timeline', buf)
        t = np.array(buf[:], dtype=np.float64)
        us = lambda i: (t[i] - t[0]) / 1965.0      # noqa: E731
        print(f'rank {rank} n_out={n_out} peers={peers}: {s.elapsed_time(e) / 200 * 1e3:.1f} us per launch | '
              f'at barrier {us(5):.2f} | passed {us(6):.2f} | pushed {us(8) if peers else 0:.2f} | '
              f'flags seen {us(9) if peers else 0:.2f} | done {us(7):.2f}', flush=True)
    del region
dist.barrier()
torch.cuda.synchronize()
print(f'rank {rank}: destroying the process group', flush=True)
dist.destroy_process_group()
print(f'rank {rank}: done', flush=True)
