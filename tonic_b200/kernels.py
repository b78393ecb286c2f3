"""Thin Python layer over the C ABI: owns device buffers (torch CUDA tensors are
used only as memory + stream plumbing) and enqueues the sm_100a kernels.

Nothing here computes with torch ops on the hot path, and nothing falls back to
the CPU: every function ends in a `_lib.call(...)` into libtonic_b200.so.
"""

import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import ptr

F32 = torch.float32


def device():
    if not torch.cuda.is_available():
        raise _lib.TonicB200Error('tonic_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def stream():
    return torch.cuda.current_stream().cuda_stream


# bytes moved across PCIe by the product's own API calls (bench.py's e2e accounting)
transfers = {'h2d': 0, 'd2h': 0}
# algorithmic FLOPs enqueued per entry point (bench.py's roofline accounting)
flops = {}
_pinned = {}


_RING = 8        # pinned staging buffers per (shape, dtype): a copy out of one may still be in flight


def to_device(x, dtype=F32, out=None):
    """numpy / torch (host or device) -> contiguous device tensor of `dtype` (`out`: write into
    this device tensor instead of a new one).  Host arrays are staged through a small ring of
    cached pinned buffers; a buffer is reused only after the copy out of it has completed (its
    own event, not a stream-wide synchronisation)."""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        x = x.to(dtype=dtype).contiguous()
        if out is None:
            return x
        out.copy_(x)
        return out
    arr = x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    target = {torch.float32: np.float32, torch.float64: np.float64,
              torch.int64: np.int64, torch.int32: np.int32}[dtype]
    arr = np.ascontiguousarray(arr, dtype=target)
    key = (arr.shape, arr.dtype.str)
    ring = _pinned.get(key)
    if ring is None:
        ring = _pinned[key] = dict(next=0, slots=[])
    if len(ring['slots']) < _RING:
        ring['slots'].append((torch.empty(arr.shape, dtype=dtype).pin_memory(), torch.cuda.Event()))
        stage, event = ring['slots'][-1]
    else:
        stage, event = ring['slots'][ring['next']]
        ring['next'] = (ring['next'] + 1) % _RING
        event.synchronize()
    if out is None:
        out = torch.empty(arr.shape, dtype=dtype, device=device())
    stage.numpy()[...] = arr
    out.copy_(stage, non_blocking=True)
    event.record()
    transfers['h2d'] += arr.nbytes
    return out


_host_stage = {}


def to_host(*tensors):
    """Device tensor(s) -> fresh numpy array(s): asynchronous copies into cached pinned buffers,
    ONE synchronisation for the whole group, then a host copy (the caller owns the result, like
    the fresh arrays the reference returns: environments/distributed.py:52-58)."""
    stages = []
    for i, t in enumerate(tensors):
        key = (i, tuple(t.shape), t.dtype)       # one staging buffer per position in the group
        stage = _host_stage.get(key)
        if stage is None:
            if len(_host_stage) > 64:       # shapes that vary (episode logs): keep the cache small
                _host_stage.clear()
            stage = _host_stage[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        stage.copy_(t.detach(), non_blocking=True)
        transfers['d2h'] += t.numel() * t.element_size()
        stages.append(stage)
    torch.cuda.current_stream().synchronize()
    out = [st.numpy().copy() for st in stages]
    return out[0] if len(out) == 1 else out


class HostBridge:
    """Fixed pinned-host and device staging buffers for ONE call of the reference-facing protocol
    with numpy arrays (agent.step / environment.step / agent.update): because the addresses never
    change, the copies and the kernels of the call can be captured into one CUDA graph
    (tonic_b200/graphs.py) -- one launch and at most one synchronisation per protocol call
    instead of a dozen small copies and launches."""

    def __init__(self):
        self.pinned, self.dev = {}, {}

    def buffers(self, name, shape, dtype=F32):
        key = (name, tuple(shape), dtype)
        if key not in self.pinned:
            self.pinned[key] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
            self.dev[key] = torch.empty(tuple(shape), dtype=dtype, device=device())
        return self.pinned[key], self.dev[key]

    def load(self, name, array, dtype=F32):
        """numpy -> pinned staging (host copy, converts bool / float64); returns (pinned, device)."""
        array = np.asarray(array)
        pinned, dev = self.buffers(name, array.shape, dtype)
        pinned.numpy()[...] = array
        transfers['h2d'] += pinned.numel() * pinned.element_size()
        return pinned, dev

    @staticmethod
    def read(pinned, dtype=None):
        """pinned staging -> fresh numpy array owned by the caller (after the synchronisation)."""
        transfers['d2h'] += pinned.numel() * pinned.element_size()
        out = pinned.numpy()
        return out.astype(dtype) if dtype is not None else out.copy()


def _count_flops(name, value):
    flops[name] = flops.get(name, 0.0) + value


def round_up(x, m):
    return (x + m - 1) // m * m


# ---------------------------------------------------------------------------
# MLP
# ---------------------------------------------------------------------------

class MlpLayout:
    """Flat float32 parameter layout of a 2-hidden-layer MLP with a linear head.

    Order: W1 [H, d_in] | b1 [H] | W2 [H, H] | b2 [H] | W3 [n_out, H] | b3 [n_out]
    | extras (e.g. log_scale [A]); every tensor starts at a multiple of 4 floats.
    """

    def __init__(self, d_in, hidden, n_out, activation, extras=()):
        self.d_in, self.hidden, self.n_out = int(d_in), int(hidden), int(n_out)
        self.act = {'tanh': _lib.ACT_TANH, 'relu': _lib.ACT_RELU}[activation]
        H = self.hidden
        off = 0
        self.offsets = {}
        for name, size in [('w1', H * d_in), ('b1', H), ('w2', H * H), ('b2', H),
                           ('w3', n_out * H), ('b3', n_out)] + \
                          [(n, s) for n, s in extras]:
            self.offsets[name] = (off, size)
            off = round_up(off + size, 4)
        self.n_params = off
        self.n_extra = sum(int(sz) for _, sz in extras)
        self.off_w1t = 0
        self.off_w2t = round_up(d_in * H, 4)
        self.n_packed = self.off_w2t + H * H
        # tf32 splits of W2 and W2^T for the tensor-core path (256-wide layers only);
        # 256-float (1 KB) aligned so they can be TMA sources
        self.tensor_core = H == 256
        split = [0, 0, 0, 0]
        if self.tensor_core:
            base = round_up(self.n_packed, 256)
            split = [base + i * H * H for i in range(4)]
            self.n_packed = base + 4 * H * H
        # W1 as the layer-1 operand of the fused forward kernel (csrc/tc_mlp.cu): image of
        # the swizzled [256 x 32] shared-memory tile, hi / lo parts
        image = [0, 0]
        self.fused_forward = self.tensor_core and d_in <= 32
        if self.fused_forward:
            image = [self.n_packed, self.n_packed + H * 32]
            self.n_packed += 2 * H * 32
        self.ldx = round_up(d_in + 1, 4)
        o = self.offsets
        self.shape = _lib.TbMlpShape(
            d_in=d_in, hidden=H, n_out=n_out, act=self.act,
            off_w1=o['w1'][0], off_b1=o['b1'][0], off_w2=o['w2'][0], off_b2=o['b2'][0],
            off_w3=o['w3'][0], off_b3=o['b3'][0], n_params=self.n_params,
            off_w1t=self.off_w1t, off_w2t=self.off_w2t, n_packed=self.n_packed,
            off_w2_hi=split[0], off_w2_lo=split[1], off_w2t_hi=split[2], off_w2t_lo=split[3],
            off_w1_img_hi=image[0], off_w1_img_lo=image[1])


class MlpInput:
    """Describes how the network input is assembled (see TbMlpInput)."""

    def __init__(self, x1, mean=None, std=None, x2=None, gather2=False, idx=None):
        self.tensors = (x1, mean, std, x2, idx)   # keep alive
        self.struct = _lib.TbMlpInput(
            d_x1=ptr(x1), dim1=x1.shape[-1], d_mean=ptr(mean), d_std=ptr(std),
            d_x2=ptr(x2), dim2=0 if x2 is None else x2.shape[-1], gather2=int(gather2),
            d_idx=ptr(idx))


_FUSED_FWD = os.environ.get('TONIC_B200_FUSED_FWD', '1') != '0'
_FUSED_BWD = os.environ.get('TONIC_B200_FUSED_BWD', '1') != '0'
_FUSED_WGRAD = os.environ.get('TONIC_B200_FUSED_WGRAD', '1') != '0'
_PEER_FUSED = os.environ.get('TONIC_B200_PEER_FUSED', '1') != '0'
_FUSED_ADAM = os.environ.get('TONIC_B200_FUSED_ADAM', '1') != '0'
_PLAIN_ACTS = os.environ.get('TONIC_B200_PLAIN_ACTS', '1') != '0'
_FUSED_TRAIN = os.environ.get('TONIC_B200_FUSED_TRAIN', '1') != '0'


class DeviceMlp:
    """Device-resident parameters (+ packed transposes) and activation workspaces."""

    def __init__(self, layout):
        self.layout = layout
        dev = device()
        self.params = torch.zeros(layout.n_params, dtype=F32, device=dev)
        self.packed = torch.zeros(layout.n_packed, dtype=F32, device=dev)
        self._ws_rows = 0
        self._gpart = None

    # -- parameter views --------------------------------------------------
    def view(self, name, shape):
        off, size = self.layout.offsets[name]
        assert int(np.prod(shape)) == size, (name, shape, size)
        return self.params[off:off + size].view(*shape)

    def pack(self):
        _lib.call('tb_mlp_pack', ctypes.byref(self.layout.shape), ptr(self.params),
                  ptr(self.packed), stream())

    # -- workspaces ---------------------------------------------------------
    def passes(self):
        """0 = FFMA kernels; 3 / 1 = tensor-core path (3xTF32 / TF32)."""
        from . import config
        if not self.layout.tensor_core or config.gemm == 'ffma':
            return 0
        if config.gemm not in ('tf32x3', 'tf32'):
            raise ValueError(f'unknown config.gemm {config.gemm!r}')
        return 3 if config.gemm == 'tf32x3' else 1

    def workspace(self, rows):
        if rows > self._ws_rows:
            L, dev = self.layout, device()
            new = lambda cols: torch.empty(rows, cols, dtype=F32, device=dev)   # noqa: E731
            self.xin = new(L.ldx)
            self.h1, self.h2, self.dz1, self.dz2 = (new(L.hidden) for _ in range(4))
            if L.tensor_core:       # low parts of the tf32 splits (h1 / dz2 hold the high parts)
                self.h1_lo, self.dz2_lo = new(L.hidden), new(L.hidden)
            self._ws_rows = rows

    def peer_region(self):
        if getattr(self, '_peer_region', None) is None:
            from . import distributed
            self._peer_region = distributed.PeerRegion(self.layout.n_params)
        return self._peer_region

    def peer_region_fused(self):
        """Symmetric region of the exchange inside the fused weight-gradient kernel (collective on
        first use, like `peer_region`)."""
        if getattr(self, '_peer_region_fused', None) is None:
            from . import distributed
            self._peer_region_fused = distributed.PeerRegion(self.layout.n_params, fused=True)
        return self._peer_region_fused

    def flat_grad(self):
        if getattr(self, '_flat_grad', None) is None:
            self._flat_grad = torch.zeros(self.layout.n_params, dtype=F32, device=device())
        return self._flat_grad

    def gpart(self, n_split):
        if self._gpart is None or self._gpart.shape[0] < n_split:
            self._gpart = torch.zeros(n_split, self.layout.n_params, dtype=F32, device=device())
        return self._gpart

    def plain_activations(self):
        """h1 / dz2 kept as ONE float32 array each (instead of tf32 hi / lo pairs): the fused
        forward / backward kernels write them straight from registers and the fused
        weight-gradient kernel splits the tiles in shared memory.  Needs all three fused kernels
        and the 3-pass mode."""
        L = self.layout
        return (_PLAIN_ACTS and self.passes() == 3 and _FUSED_FWD and _FUSED_BWD and L.fused_forward
                and L.n_out <= 8 and self.fused_wgrad(L.n_extra))

    def fused_wgrad(self, n_extra=0):
        """All weight gradients in one launch, reduced in the kernel to ONE flat gradient
        (csrc/tc_gemm.cu::tc_wgrad_all_kernel)."""
        L = self.layout
        return (bool(self.passes()) and _FUSED_WGRAD and L.d_in + 1 <= 32 and L.n_out <= 8
                and L.n_out + n_extra <= 12)

    def splits_for(self, rows, n_extra=0):
        """Row splits of the weight-gradient partial sums.  Tensor-core path: 2 column tiles x
        splits CTAs; the fused kernel needs all of them resident (<= 148).  Unfused chain: the
        narrow gradients use twice the splits of the W2 kernel (see `w2_splits`)."""
        from . import config
        if self.fused_wgrad(n_extra):
            return max(1, min(config.wgrad_splits_tc, rows // 32))
        if self.passes():
            return max(1, min(2 * config.wgrad_splits_tc, rows // 32))
        return max(1, min(config.wgrad_splits, rows // 128))

    def w2_splits(self, n_split):
        """Splits of the W2 block inside `n_split` partial slots (0 = same as the rest)."""
        return max(1, n_split // 2) if self.passes() else 0

    def w2_range(self):
        """Parameters whose partial gradients come from the tensor-core weight-gradient kernel
        (its own split count): W2 and, right behind it in the layout, b2."""
        off, size = self.layout.offsets['w2']
        if self.passes():
            assert self.layout.offsets['b2'][0] == off + size
            size += self.layout.hidden
        return off, off + size

    # -- kernels ------------------------------------------------------------
    def forward(self, inp, rows, out, save=False, skip=None, params=None, packed=None, vloss=None):
        """out[rows, n_out] = head pre-activations. `params`/`packed` override the
        parameter set (target networks share the layout).  `vloss` = (targets, idx, dout, stats):
        squared-error loss of a single-output head; `self.vloss_fused` tells the caller whether
        the forward kernel computed it (otherwise kernels.mse_loss has to be launched)."""
        self.vloss_fused = False
        passes = self.passes()
        L = self.layout
        # one-kernel tensor-core forward (csrc/tc_mlp.cu): no activation round trip through
        # global memory unless a backward pass follows
        fused = bool(passes) and L.fused_forward and L.n_out <= 8 and _FUSED_FWD
        if save or (passes and not fused):
            self.workspace(rows)
        params = self.params if params is None else params
        packed = self.packed if packed is None else packed
        flops = 2.0 * rows * (L.d_in * L.hidden + L.hidden * L.hidden + L.hidden * L.n_out)
        if passes:
            _count_flops('tb_mlp_forward_tc', flops)
            _count_flops('tb_tc_mlp_forward' if fused else 'tb_tc_gemm256_fwd',
                         flops if fused else 2.0 * rows * L.hidden * L.hidden)
            if fused and vloss is not None and L.n_out == 1:
                targets, idx, dout, stats = vloss
                self.vloss_fused = True
                _lib.call('tb_tc_mlp_forward_vloss', ctypes.byref(L.shape), ptr(params), ptr(packed),
                          ctypes.byref(inp.struct), rows, ptr(out), ptr(self.xin) if save else None,
                          *((None, None, None) if not save
                            else (ptr(self.h1), None if self.plain_activations() else ptr(self.h1_lo),
                                  ptr(self.h2))),
                          passes, ptr(targets), ptr(idx), ptr(dout),
                          dout.shape[-1] if dout.dim() > 1 else 1, ptr(stats), _lib.STAT_VALUE, 1,
                          ptr(skip), stream())
                return out
            _lib.call('tb_mlp_forward_tc', ctypes.byref(L.shape), ptr(params), ptr(packed),
                      ctypes.byref(inp.struct), rows, ptr(out), ptr(self.xin) if save else None,
                      *((None, None, None) if fused and not save
                        else (ptr(self.h1), None if fused and self.plain_activations() else ptr(self.h1_lo),
                              ptr(self.h2))),
                      passes, ptr(skip), stream())
            return out
        _count_flops('tb_mlp_forward', flops)
        _lib.call('tb_mlp_forward', ctypes.byref(L.shape), ptr(params), ptr(packed),
                  ctypes.byref(inp.struct), rows, ptr(out),
                  ptr(self.xin) if save else None, ptr(self.h1) if save else None,
                  ptr(self.h2) if save else None, ptr(skip), stream())
        return out

    def fused_train(self):
        """forward -> loss -> backward as ONE launch (csrc/tc_mlp.cu::tc_mlp_train_kernel)."""
        L = self.layout
        return (bool(self.passes()) and _FUSED_TRAIN and _FUSED_FWD and _FUSED_BWD and L.fused_forward
                and L.n_out <= 8)

    def train_step(self, inp, rows, dout, stats, idx=None, targets=None, policy=None, out=None,
                   skip=None):
        """Value regression (`targets`) or Gaussian policy loss (`policy` = dict(log_scale, actions,
        advantages, log_probs, ratio_clip, entropy_coeff)) on the minibatch `inp`: activations for
        the weight-gradient kernel end up in the workspaces, the head gradient in `dout`."""
        L = self.layout
        self.workspace(rows)
        plain = self.plain_activations()
        flops = 2.0 * rows * (L.d_in * L.hidden + L.hidden * L.hidden + L.hidden * L.n_out) \
            + 2.0 * rows * L.hidden * (L.n_out + L.hidden)
        _count_flops('tb_tc_mlp_train', flops)
        pol = policy or {}
        _lib.call('tb_tc_mlp_train', ctypes.byref(L.shape), ptr(self.params), ptr(self.packed),
                  ctypes.byref(inp.struct), rows, 0 if policy is None else 1, ptr(idx), ptr(targets),
                  ptr(pol.get('log_scale')), ptr(pol.get('actions')), ptr(pol.get('advantages')),
                  ptr(pol.get('log_probs')), float(pol.get('ratio_clip', 0.0)),
                  float(pol.get('entropy_coeff', 0.0)), ptr(stats), ptr(out), ptr(self.xin),
                  ptr(self.h1), None if plain else ptr(self.h1_lo), ptr(self.h2), ptr(dout),
                  dout.shape[-1] if dout.dim() > 1 else 1, ptr(self.dz2),
                  None if plain else ptr(self.dz2_lo), ptr(self.dz1), self.passes(), ptr(skip), stream())

    def backward(self, dout, rows, dx=None, dx_col0=0, skip=None, params=None, packed=None):
        L = self.layout
        params = self.params if params is None else params
        flops = 2.0 * rows * L.hidden * (L.n_out + L.hidden + (0 if dx is None else dx.shape[-1]))
        passes = self.passes()
        if passes:
            _count_flops('tb_mlp_backward_tc', flops)
            if L.n_out <= 8 and _FUSED_BWD:      # one kernel: head gradient + hidden-layer GEMM
                _count_flops('tb_tc_mlp_backward', 2.0 * rows * L.hidden * (L.n_out + L.hidden))
            else:
                _count_flops('tb_tc_gemm256_bwd', 2.0 * rows * L.hidden * L.hidden)
            _lib.call('tb_mlp_backward_tc', ctypes.byref(L.shape), ptr(params),
                      ptr(self.packed if packed is None else packed), ptr(dout), dout.shape[-1],
                      ptr(self.h1), None if self.plain_activations() else ptr(self.h1_lo), ptr(self.h2),
                      rows, ptr(self.dz2), None if self.plain_activations() else ptr(self.dz2_lo),
                      ptr(self.dz1), ptr(dx), dx_col0,
                      0 if dx is None else dx.shape[-1], passes, ptr(skip), stream())
            return
        _count_flops('tb_mlp_backward', flops)
        _lib.call('tb_mlp_backward', ctypes.byref(L.shape), ptr(params), ptr(dout), dout.shape[-1],
                  ptr(self.h1), ptr(self.h2), rows, ptr(self.dz2), ptr(self.dz1), ptr(dx),
                  dx_col0, 0 if dx is None else dx.shape[-1], ptr(skip), stream())

    def wgrad(self, dout, rows, n_split, n_extra=0, off_extra=0, skip=None, fuse=None):
        """`fuse` = (adam, grad_scale, stats, kl_threshold, stop, reduce_stats): the fused kernel also
        runs the optimizer step (no clipping) and, with several ranks, the gradient exchange over
        NVLink peer memory before it; `self.applied` tells the caller."""
        self.applied = False
        gpart = self.gpart(n_split)
        L = self.layout
        flops = 2.0 * rows * (L.hidden * L.hidden + L.hidden * (L.d_in + 2)
                              + (L.n_out + n_extra) * (L.hidden + 1))
        passes = self.passes()
        if self.fused_wgrad(n_extra):
            # one launch: tensor-core dW2 + FFMA narrow gradients + in-kernel reduction; the
            # result is the flat gradient (n_split = 1 for the optimizer / the exchange)
            _count_flops('tb_mlp_wgrad_fused', flops)
            if getattr(self, '_wgrad_sync', None) is None:
                self._wgrad_sync = torch.zeros(1, dtype=torch.int64, device=device())
            flat = self.flat_grad()
            opt = packed = stats = stop = peers = epoch = reduce_stats = None
            scale, kl = 0.0, -1.0
            if fuse is not None:
                adam, scale, stats, kl, stop, reduce_stats = fuse
                opt, packed = ctypes.byref(adam.struct), self.packed
                from . import distributed
                if distributed.world() > 1:      # gradient exchange inside the same launch
                    region = self.peer_region_fused()
                    peers, epoch = ctypes.byref(region.struct), region.epoch
                    reduce_stats = stats if reduce_stats is None else reduce_stats
            plain = self.plain_activations()
            _lib.call('tb_mlp_wgrad_fused', ctypes.byref(L.shape), ptr(self.xin), ptr(self.h1),
                      None if plain else ptr(self.h1_lo), ptr(self.h2), ptr(self.dz1), ptr(self.dz2),
                      None if plain else ptr(self.dz2_lo), ptr(dout), dout.shape[-1], n_extra, off_extra, rows,
                      ptr(gpart), n_split, ptr(flat), ptr(self._wgrad_sync), passes, opt,
                      ptr(packed), scale, ptr(stats), kl, ptr(stop), ptr(skip), peers, ptr(epoch),
                      ptr(reduce_stats), stream())
            self.reduced = True
            self.applied = fuse is not None
            return flat
        self.reduced = False
        if self.plain_activations():
            raise _lib.TonicB200Error('plain activations were saved but the fused weight-gradient '
                                      f'kernel does not cover n_extra={n_extra}')
        if passes:
            _count_flops('tb_mlp_wgrad_tc', flops)
            _count_flops('tb_tc_wgrad256', 2.0 * rows * L.hidden * L.hidden)
            _lib.call('tb_mlp_wgrad_tc', ctypes.byref(L.shape), ptr(self.xin), ptr(self.h1),
                      ptr(self.h1_lo), ptr(self.h2), ptr(self.dz1), ptr(self.dz2),
                      ptr(self.dz2_lo), ptr(dout), dout.shape[-1], n_extra, off_extra, rows,
                      ptr(gpart), n_split, self.w2_splits(n_split), passes, ptr(skip), stream())
            return gpart
        _count_flops('tb_mlp_wgrad', flops)
        _lib.call('tb_mlp_wgrad', ctypes.byref(L.shape), ptr(self.xin), ptr(self.h1),
                  ptr(self.h2), ptr(self.dz1), ptr(self.dz2), ptr(dout), dout.shape[-1],
                  n_extra, off_extra, rows, ptr(gpart), n_split, ptr(skip), stream())
        return gpart


class Adam:
    """torch.optim.Adam state for one flat parameter buffer (device resident)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        dev = device()
        self.params = params
        self.m = torch.zeros_like(params)
        self.v = torch.zeros_like(params)
        self.step_count = torch.zeros(2, dtype=torch.int32, device=dev)
        self.struct = _lib.TbAdam(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps,
                                  n_params=params.numel(), d_params=ptr(params),
                                  d_m=ptr(self.m), d_v=ptr(self.v), d_step=ptr(self.step_count))

    def step(self, mlp, gpart, n_split, grad_scale, skip=None, stats=None,
             kl_threshold=-1.0, stop=None, n_split_w2=0):
        _lib.call('tb_adam_step', ctypes.byref(self.struct), ctypes.byref(mlp.layout.shape),
                  ptr(mlp.packed), ptr(gpart), n_split, n_split_w2, grad_scale, ptr(skip),
                  ptr(stats), kl_threshold, ptr(stop), stream())


class GradientClipper:
    """torch.nn.utils.clip_grad_norm_ over ALL variables of one updater's optimizer
    (reference: updaters/actors.py:37-38,96-98; critics.py:24-25,177-178 -- the twin-critic
    updaters clip the joint norm of both critics).  `add` turns one network's split partial
    sums into its flat (all-reduced) gradient and accumulates the squared norm on the
    device; `finish` scales every pending gradient by min(1, max_norm / (norm + 1e-6)) and
    runs the Adam steps.  `deferred` = the caller adds several networks before `finish`."""

    def __init__(self, max_norm):
        self.max_norm = float(max_norm)
        self.sumsq = None
        self.pending = []
        self.deferred = False

    def add(self, adam, mlp, gpart, n_split, rows_local, rows_global, skip, stats, kl_threshold,
            stop, reduce_stats):
        from . import distributed
        if self.sumsq is None:
            self.sumsq = torch.zeros(1, dtype=torch.float64, device=device())
        if not self.pending:
            self.sumsq.zero_()
        n = mlp.layout.n_params
        flat = mlp.flat_grad()
        if rows_local > 0 and gpart.data_ptr() != flat.data_ptr():
            w2_lo, w2_hi = mlp.w2_range()
            _lib.call('tb_reduce_partials', ptr(gpart), n_split,
                      mlp.w2_splits(n_split) if n_split > 1 else 0, w2_lo,
                      w2_hi, n, ptr(flat), None, stream())
        elif rows_local == 0:
            flat.zero_()
        if distributed.world() > 1:
            distributed.all_reduce(flat)
            reduce_stats = stats if reduce_stats is None else reduce_stats
            if reduce_stats is not None:
                distributed.all_reduce(reduce_stats)
        _lib.call('tb_grad_sqnorm', ptr(flat), n, ptr(self.sumsq), ptr(skip), stream())
        self.pending.append((adam, mlp, flat, 1.0 / rows_global, skip, stats, kl_threshold, stop))

    def finish(self):
        for adam, mlp, flat, scale, skip, stats, kl_threshold, stop in self.pending:
            _lib.call('tb_grad_clip', ptr(flat), mlp.layout.n_params, ptr(self.sumsq), scale,
                      self.max_norm, ptr(skip), stream())
            adam.step(mlp, flat, 1, scale, skip=skip, stats=stats, kl_threshold=kl_threshold,
                      stop=stop)
        self.pending.clear()


def make_clipper(gradient_clip):
    return GradientClipper(gradient_clip) if gradient_clip and gradient_clip > 0 else None


def wgrad_and_apply(adam, mlp, dout, rows, rows_global, n_extra=0, off_extra=0, skip=None,
                    stats=None, kl_threshold=-1.0, stop=None, reduce_stats=None, clip=None):
    """Weight gradients of the minibatch whose activations `mlp` holds, then the optimizer step
    (loss.backward() ... optimizer.step(), e.g. updaters/critics.py:23-26).  Without gradient
    clipping on the tensor-core path: ONE launch (weight gradients, in-kernel reduction, the
    exchange between ranks when every rank owns rows of every minibatch, and Adam).  Otherwise weight gradients -> [clip] -> [exchange] -> Adam (`apply_gradients`)."""
    from . import distributed
    n_split = mlp.splits_for(rows, n_extra)
    gpart = None
    if rows > 0:
        fuse = None
        if (_FUSED_ADAM and clip is None and mlp.fused_wgrad(n_extra) and
                (distributed.world() == 1 or balanced_exchange())):
            fuse = (adam, 1.0 / rows_global, stats, kl_threshold, stop, reduce_stats)
        gpart = mlp.wgrad(dout, rows, n_split, n_extra=n_extra, off_extra=off_extra, skip=skip,
                          fuse=fuse)
        if mlp.applied:
            return
    apply_gradients(adam, mlp, gpart, n_split, rows, rows_global, skip=skip, stats=stats,
                    kl_threshold=kl_threshold, stop=stop, reduce_stats=reduce_stats, clip=clip)


def apply_gradients(adam, mlp, gpart, n_split, rows_local, rows_global, skip=None, stats=None,
                    kl_threshold=-1.0, stop=None, reduce_stats=None, clip=None):
    """Adam step from the split weight gradients of this rank.  Single process:
    the partial sums are reduced inside the Adam kernel.  Several ranks: partial
    sums -> flat gradient -> all-reduce (and the statistics block) -> Adam with
    1 / (global rows), which equals the single-process mean.
    `stats` steers the Adam kernel (policy updaters: skip when every advantage is
    zero, KL early stop); `reduce_stats` is a statistics block that only needs the
    sum over ranks (defaults to `stats`)."""
    from . import distributed
    if gpart is not None and getattr(mlp, 'reduced', False) and rows_local > 0:
        n_split = 1         # `gpart` is the flat gradient the fused weight-gradient kernel reduced
    if clip is not None:
        clip.add(adam, mlp, gpart, n_split, rows_local, rows_global, skip, stats, kl_threshold,
                 stop, reduce_stats)
        if not clip.deferred:
            clip.finish()
        return
    w2_splits = mlp.w2_splits(n_split) if n_split > 1 else 0
    w2_lo, w2_hi = mlp.w2_range()
    if distributed.world() == 1:
        adam.step(mlp, gpart, n_split, 1.0 / rows_global, skip=skip, stats=stats,
                  kl_threshold=kl_threshold, stop=stop, n_split_w2=w2_splits)
        return
    if config_peer_reduce():
        # fused path: publish flat gradient + statistics to the NVLink-mapped region, then one
        # kernel sums every rank's slot with peer loads and applies Adam
        region = mlp.peer_region()
        reduce_stats = stats if reduce_stats is None else reduce_stats
        _lib.call('tb_peer_publish', ctypes.byref(region.struct),
                  ptr(gpart) if rows_local > 0 else None, n_split, w2_splits, w2_lo, w2_hi,
                  mlp.layout.n_params, ptr(reduce_stats), ptr(region.epoch), ptr(region.block_counter), ptr(skip),
                  stream())
        _lib.call('tb_adam_step_peers', ctypes.byref(adam.struct), ctypes.byref(mlp.layout.shape),
                  ptr(mlp.packed), ctypes.byref(region.struct), 1.0 / rows_global,
                  ptr(region.epoch), ptr(skip), ptr(reduce_stats), int(stats is not None),
                  kl_threshold, ptr(stop), stream())
        return
    flat = mlp.flat_grad()
    if rows_local > 0 and gpart.data_ptr() != flat.data_ptr():
        _lib.call('tb_reduce_partials', ptr(gpart), n_split, w2_splits, w2_lo, w2_hi,
                  mlp.layout.n_params, ptr(flat), None, stream())
    elif rows_local == 0:
        flat.zero_()
    distributed.all_reduce(flat)
    reduce_stats = stats if reduce_stats is None else reduce_stats
    if reduce_stats is not None:
        distributed.all_reduce(reduce_stats)
    adam.step(mlp, flat, 1, 1.0 / rows_global, skip=skip, stats=stats,
              kl_threshold=kl_threshold, stop=stop)


def balanced_exchange():
    """Several ranks: may the gradient exchange run inside the fused weight-gradient kernel?  Every
    rank must then launch that kernel for every minibatch, i.e. own rows of every minibatch: true
    for the device permutations (each rank contributes batch_size / world rows), not for the
    reference-exact global permutation of the parity mode, where a rank's share of a minibatch can
    be empty -- that mode keeps the publish / pull kernels."""
    from . import config
    return config.peer_reduce and config.indices == 'device' and _PEER_FUSED


def config_peer_reduce():
    from . import config
    return config.peer_reduce


def soft_update(target, online, tau):
    _lib.call('tb_soft_update', ptr(target), ptr(online), target.numel(), tau, stream())


# ---------------------------------------------------------------------------
# heads / losses
# ---------------------------------------------------------------------------

def gauss_sample(loc_pre, log_scale, actions, log_probs, eps=None, seed=0, counter=0,
                 device_counter=None):
    rows, act = loc_pre.shape
    _lib.call('tb_gauss_sample', ptr(loc_pre), ptr(log_scale), ptr(eps), seed, counter,
              ptr(device_counter), rows, act, ptr(actions), ptr(log_probs), stream())


def rollout_fused(env_struct, mlp, log_scale, norm_mean, norm_std, T, buffers, env_obs,
                  moment_sums, seed, counter, counter_stride, device_counter):
    """All T vector steps of the on-policy collector in one launch
    (csrc/mlp.cu::rollout_kernel; reference loop utils/trainer.py:44-50)."""
    b = buffers
    _lib.call('tb_rollout_fused', ctypes.byref(env_struct), ctypes.byref(mlp.layout.shape),
              ptr(mlp.params), ptr(mlp.packed), ptr(log_scale), ptr(norm_mean), ptr(norm_std),
              int(T), ptr(b['observations']), ptr(b['actions']), ptr(b['next_observations']),
              ptr(b['rewards']), ptr(b['resets']), ptr(b['terminations']), ptr(b['log_probs']),
              ptr(env_obs), ptr(moment_sums), int(seed), int(counter), int(counter_stride),
              ptr(device_counter), stream())


def act_env_step(env_struct, loc_pre, log_scale, seed, counter, device_counter, actions, log_probs,
                 moment_sums, observations, next_observations, rewards, resets, terminations):
    """Sample + log-prob + normaliser record + environment step of one vector step in one launch
    (csrc/env_step.cu::act_env_step_kernel; trainer.py:44-50 after the actor forward)."""
    _lib.call('tb_act_env_step', ctypes.byref(env_struct), ptr(loc_pre), ptr(log_scale), int(seed),
              int(counter), ptr(device_counter), ptr(actions), ptr(log_probs), ptr(moment_sums),
              ptr(observations), ptr(next_observations), ptr(rewards), ptr(resets), ptr(terminations),
              stream())


def counter_add(counter, delta):
    _lib.call('tb_counter_add', ptr(counter), int(delta), stream())


def new_counter():
    return torch.zeros(1, dtype=torch.int64, device=device())


def gauss_policy_loss(loc_pre, log_scale, actions, advantages, old_log_probs, idx, rows, dout,
                      stats, ratio_clip, entropy_coeff, skip=None):
    _lib.call('tb_gauss_policy_loss', ptr(loc_pre), ptr(log_scale), ptr(actions),
              ptr(advantages), ptr(old_log_probs), ptr(idx), rows, actions.shape[-1],
              ratio_clip, entropy_coeff, ptr(dout), ptr(stats), ptr(skip), stream())


def mse_loss(values, targets, idx, rows, dout, stats, stat_slot=_lib.STAT_VALUE,
             count_rows=True, skip=None):
    _lib.call('tb_mse_loss', ptr(values), ptr(targets), ptr(idx), rows, ptr(dout),
              dout.shape[-1] if dout.dim() > 1 else 1, ptr(stats), stat_slot, int(count_rows),
              ptr(skip), stream())


# ---------------------------------------------------------------------------
# replay helpers / normaliser
# ---------------------------------------------------------------------------

def lambda_returns(values, next_values, rewards, resets, terminations, returns,
                   discount_factor, trace_decay):
    T, N = rewards.shape
    _lib.call('tb_lambda_returns', ptr(values), ptr(next_values), ptr(rewards), ptr(resets),
              ptr(terminations), ptr(returns), T, N, discount_factor, trace_decay, stream())


def advantages(returns, values, out, workspace, n_global=None, phase=0):
    n = returns.numel()
    _lib.call('tb_advantages', ptr(returns), ptr(values), ptr(out), n, ptr(workspace),
              n if n_global is None else n_global, phase, stream())


def moments_record(x, sums):
    rows, dim = x.shape
    _lib.call('tb_moments_record', ptr(x), rows, dim, ptr(sums), stream())


def moments_update(sums, running, count, mean, std, eps=1e-2):
    _lib.call('tb_moments_update', ptr(sums), ptr(running), ptr(count), ptr(mean), ptr(std),
              mean.numel(), eps, stream())


class ArrayStats:
    """Device accumulator of (count, sum, sum of squares, min, max) of float32
    arrays (csrc/heads.cu::array_stats_kernel)."""

    _NEG_INF_ENC = np.array([-np.inf]).view(np.int64)[0] ^ np.int64(0x7fffffffffffffff)

    def __init__(self):
        self.acc = torch.zeros(5, dtype=torch.float64, device=device())
        self.reset()

    def reset(self):
        init = np.zeros(5)
        init[3] = np.inf
        init.view(np.int64)[4] = self._NEG_INF_ENC
        self.acc.copy_(torch.from_numpy(init))
        self.items = 0

    def add(self, x, items=1):
        _lib.call('tb_array_stats', ptr(x), x.numel(), ptr(self.acc), stream())
        self.items += items

    def read(self):
        """-> (count, sum, sum_sq, min, max) as python floats (synchronises)."""
        host = self.acc.cpu().numpy()
        bits = host.view(np.int64)
        ext = []
        for b in (bits[3], bits[4]):
            b = b if b >= 0 else b ^ np.int64(0x7fffffffffffffff)
            ext.append(float(np.array([b], np.int64).view(np.float64)[0]))
        return float(host[0]), float(host[1]), float(host[2]), ext[0], ext[1]


def profile_begin():
    _lib.call('tb_profile_begin')


def profile_end():
    """-> {entry point: (launches, total_ms)} since profile_begin()."""
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.call('tb_profile_end', buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, count, ms = line.split()
        out[name] = (int(count), float(ms))
    return out


# ---------------------------------------------------------------------------
# off-policy heads / targets / actor losses
# ---------------------------------------------------------------------------

def tanh_action(pre, out, mode=0, noise32=None, noise64=None, seed=0, counter=0,
                noise_scale=0.0, noise_clip=float('inf')):
    rows, act = out.shape
    _lib.call('tb_tanh_action', ptr(pre), rows, act, mode, ptr(noise32), ptr(noise64), seed,
              counter, noise_scale, noise_clip, ptr(out), stream())


def squashed_sample(pre, actions, log_probs=None, eps=None, eps_out=None, seed=0, counter=0,
                    greedy=False):
    rows, act = actions.shape
    _lib.call('tb_squashed_sample', ptr(pre), ptr(eps), seed, counter, rows, act, int(greedy),
              ptr(actions), ptr(log_probs), ptr(eps_out), stream())


def q_target(rewards, terminations, idx, discount_factor, q1, q2, log_probs, entropy_coeff,
             rows, targets, discounts=None):
    """`discounts`: the replay's stored discounts column (n-step returns) instead of
    (1 - terminations) * discount_factor."""
    if discounts is not None:
        _lib.call('tb_q_target_discounts', ptr(rewards), ptr(discounts), ptr(idx), ptr(q1), ptr(q2),
                  ptr(log_probs), entropy_coeff, rows, ptr(targets), stream())
        return
    _lib.call('tb_q_target', ptr(rewards), ptr(terminations), ptr(idx), discount_factor, ptr(q1),
              ptr(q2), ptr(log_probs), entropy_coeff, rows, ptr(targets), stream())


def replay_accumulate_n_steps(rewards, discounts, next_observations, resets, index, size,
                              return_steps):
    max_size, workers = rewards.shape[:2]
    _lib.call('tb_replay_accumulate_n_steps', ptr(rewards), ptr(discounts), ptr(next_observations),
              ptr(resets), index, size, max_size, workers, next_observations.shape[-1],
              return_steps, stream())


def q_actor_loss(q1, q2, log_probs, entropy_coeff, rows, dout1, dout2, stats):
    _lib.call('tb_q_actor_loss', ptr(q1), ptr(q2), ptr(log_probs), entropy_coeff, rows,
              ptr(dout1), ptr(dout2), ptr(stats), stream())


def dpg_head_grad(dqda, actions, dout):
    rows, act = actions.shape
    _lib.call('tb_dpg_head_grad', ptr(dqda), ptr(actions), rows, act, ptr(dout), stream())


def sac_head_grad(pre, eps, actions, dqda1, dqda2, entropy_coeff, dout):
    rows, act = actions.shape
    _lib.call('tb_sac_head_grad', ptr(pre), ptr(eps), ptr(actions), ptr(dqda1), ptr(dqda2),
              entropy_coeff, rows, act, ptr(dout), stream())


def permutation(seed, stream_id, out, device_counter=None):
    """out[i] = pseudo-random bijection of [0, len(out)) (device fast-mode indices)."""
    _lib.call('tb_permutation', seed, stream_id, ptr(device_counter), out.numel(), ptr(out),
              stream())


# ---------------------------------------------------------------------------
# tensor-core GEMM (tcgen05 / TMA / TMEM)
# ---------------------------------------------------------------------------

def split_tf32(x, hi, lo):
    _lib.call('tb_split_tf32', ptr(x), ptr(hi), ptr(lo), x.numel(), stream())


def tc_gemm256(a_hi, a_lo, b_hi, b_lo, rows, out, passes=3, epilogue=2, act=0, bias=None,
               aux_hi=None, aux_lo=None, out_lo=None, head_w=None, head_b=None, head_out=None,
               skip=None):
    """out[rows, 256] = epilogue(A . B^T) on the tensor cores (see tb_tc_gemm256)."""
    _count_flops('tb_tc_gemm256', 2.0 * rows * 256 * 256)
    _lib.call('tb_tc_gemm256', ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), rows, passes, epilogue,
              act, ptr(bias), ptr(aux_hi), ptr(aux_lo), ptr(out), ptr(out_lo), ptr(head_w),
              ptr(head_b), ptr(head_out), 0 if head_w is None else head_w.shape[0], ptr(skip),
              stream())


def tc_wgrad256(dz_hi, dz_lo, h_hi, h_lo, rows, gpart, n_split, n_params, off_w2, passes=3,
                skip=None, off_b2=-1):
    """gpart[s, off_w2 + n*256 + k] = sum_m dz[m, n] h[m, k] over the rows of split s."""
    _count_flops('tb_tc_wgrad256', 2.0 * rows * 256 * 256)
    _lib.call('tb_tc_wgrad256', ptr(dz_hi), ptr(dz_lo), ptr(h_hi), ptr(h_lo), rows, passes,
              ptr(gpart), n_split, n_params, off_w2, off_b2, ptr(skip), stream())
