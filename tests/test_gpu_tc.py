"""tcgen05 tensor-core GEMM (TMA + TMEM) against float64 matmul."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def K():
    from tonic_b200 import kernels
    kernels.device()
    return kernels


def split(K, x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    K.split_tf32(x, hi, lo)
    return hi, lo


def test_split_is_exact(K):
    x = torch.randn(1000, device='cuda') * 3
    hi, lo = split(K, x)
    assert torch.equal(hi + lo, x)
    assert (hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0


@pytest.mark.parametrize('rows', [128, 100, 1000, 16384, 148 * 128 * 2 + 77])
@pytest.mark.parametrize('passes', [3, 1])
def test_tc_gemm_plain(K, rows, passes):
    g = torch.Generator().manual_seed(rows)
    a = torch.randn(rows, 256, generator=g)
    b = torch.randn(256, 256, generator=g) * 0.1
    ref = (a.double() @ b.double().T)
    a_hi, a_lo = split(K, a.cuda())
    b_hi, b_lo = split(K, b.cuda())
    out = torch.full((rows, 256), float('nan'), device='cuda')
    K.tc_gemm256(a_hi, a_lo, b_hi, b_lo, rows, out, passes=passes, epilogue=2)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # 3xTF32: fp32-grade (lo operands are truncated to 11 bits: <= 1e-5 of the output scale, typically 3e-6); plain TF32: ~1e-3
    assert err <= (1e-5 if passes == 3 else 5e-3) * scale, (err, scale)
    if passes == 1:
        assert err > 1e-6 * scale     # really ran on the tf32 path


@pytest.mark.parametrize('act', ['tanh', 'relu'])
def test_tc_gemm_epilogues(K, act):
    rows = 700
    g = torch.Generator().manual_seed(7)
    a = torch.randn(rows, 256, generator=g) * 0.3
    w = torch.randn(256, 256, generator=g) * 0.1
    bias = torch.randn(256, generator=g) * 0.1
    f = torch.tanh if act == 'tanh' else torch.relu
    act_id = 0 if act == 'tanh' else 1
    a_hi, a_lo = split(K, a.cuda())
    w_hi, w_lo = split(K, w.cuda())
    # forward: h2 = act(a W^T + b), written as a tf32 split
    out_hi = torch.empty(rows, 256, device='cuda')
    out_lo = torch.empty(rows, 256, device='cuda')
    K.tc_gemm256(a_hi, a_lo, w_hi, w_lo, rows, out_hi, passes=3, epilogue=0, act=act_id,
                 bias=bias.cuda(), out_lo=out_lo)
    ref = f(a @ w.T + bias)
    np.testing.assert_allclose((out_hi + out_lo).cpu(), ref, rtol=2e-5, atol=3e-5)
    assert (out_hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0
    # backward: dz1 = (dz2 W) * act'(h1): B operand is W^T (rows = output index)
    h1 = f(torch.randn(rows, 256, generator=g))
    h_hi, h_lo = split(K, h1.cuda())
    wt_hi, wt_lo = split(K, w.T.contiguous().cuda())
    out = torch.empty(rows, 256, device='cuda')
    K.tc_gemm256(a_hi, a_lo, wt_hi, wt_lo, rows, out, passes=3, epilogue=1, act=act_id,
                 aux_hi=h_hi, aux_lo=h_lo)
    grad = (1 - h1 * h1) if act == 'tanh' else (h1 > 0).float()
    ref = (a @ w) * grad
    np.testing.assert_allclose(out.cpu(), ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('rows,n_split', [(64, 1), (1000, 3), (16384, 74), (100, 7)])
@pytest.mark.parametrize('passes', [3, 1])
def test_tc_wgrad(K, rows, n_split, passes):
    g = torch.Generator().manual_seed(rows)
    dz = torch.randn(rows, 256, generator=g) * 0.1
    h = torch.randn(rows, 256, generator=g)
    ref = dz.double().T @ h.double()
    dz_hi, dz_lo = split(K, dz.cuda())
    h_hi, h_lo = split(K, h.cuda())
    n_params, off = 70400, 4352
    gpart = torch.full((n_split, n_params), float('nan'), device='cuda')
    K.tc_wgrad256(dz_hi, dz_lo, h_hi, h_lo, rows, gpart, n_split, n_params, off, passes=passes)
    got = gpart[:, off:off + 65536].sum(0).view(256, 256).cpu().double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= (1e-5 if passes == 3 else 5e-3) * scale, (err, scale)
    # nothing outside the W2 block is touched
    assert torch.isnan(gpart[:, :off]).all() and torch.isnan(gpart[:, off + 65536:]).all()
    # with off_b2 the same pass also produces the bias gradient (column sums of dz) through
    # N = 16 MMAs against a block of ones
    gpart.fill_(float('nan'))
    K.tc_wgrad256(dz_hi, dz_lo, h_hi, h_lo, rows, gpart, n_split, n_params, off, passes=passes,
                  off_b2=off + 65536)
    got = gpart[:, off:off + 65536].sum(0).view(256, 256).cpu().double()
    assert (got - ref).abs().max().item() <= (1e-5 if passes == 3 else 5e-3) * scale
    sums = gpart[:, off + 65536:off + 65536 + 256].sum(0).cpu().double()
    ref_b = dz.double().sum(0)
    tol = (1e-5 if passes == 3 else 2e-3) * max(1.0, ref_b.abs().max().item())
    assert (sums - ref_b).abs().max().item() <= tol
    assert torch.isnan(gpart[:, :off]).all() and torch.isnan(gpart[:, off + 65536 + 256:]).all()


def test_tc_gemm_throughput_smoke(K):
    rows = 16384
    a = torch.randn(rows, 256, device='cuda')
    b = torch.randn(256, 256, device='cuda') * 0.1
    a_hi, a_lo = split(K, a)
    b_hi, b_lo = split(K, b)
    out = torch.empty(rows, 256, device='cuda')
    for passes in (3, 1):
        for _ in range(3):
            K.tc_gemm256(a_hi, a_lo, b_hi, b_lo, rows, out, passes=passes)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(20):
            K.tc_gemm256(a_hi, a_lo, b_hi, b_lo, rows, out, passes=passes)
        end.record()
        torch.cuda.synchronize()
        us = start.elapsed_time(end) / 20 * 1e3
        print(f'tc_gemm256 rows={rows} passes={passes}: {us:.1f} us, '
              f'{2 * rows * 65536 / us / 1e6:.1f} TFLOP/s (fp32-equivalent)')
        gpart = torch.empty(74, 70000, device='cuda')
        for _ in range(3):
            K.tc_wgrad256(a_hi, a_lo, a_hi, a_lo, rows, gpart, 74, 70000, 0, passes=passes)
        start.record()
        for _ in range(20):
            K.tc_wgrad256(a_hi, a_lo, a_hi, a_lo, rows, gpart, 74, 70000, 0, passes=passes)
        end.record()
        torch.cuda.synchronize()
        us = start.elapsed_time(end) / 20 * 1e3
        print(f'tc_wgrad256 rows={rows} passes={passes}: {us:.1f} us, '
              f'{2 * rows * 65536 / us / 1e6:.1f} TFLOP/s (fp32-equivalent)')


# ---- one-kernel forward pass (csrc/tc_mlp.cu) --------------------------------------------
def _reference_forward(net, x, act):
    L, H = net.layout, net.layout.hidden
    p = {k: net.view(k, s).cpu().double() for k, s in
         dict(w1=(H, L.d_in), b1=(H,), w2=(H, H), b2=(H,), w3=(L.n_out, H), b3=(L.n_out,)).items()}
    f = torch.tanh if act == 'tanh' else torch.relu
    h1 = f(x.double() @ p['w1'].T + p['b1'])
    h2 = f(h1 @ p['w2'].T + p['b2'])
    return h1, h2, h2 @ p['w3'].T + p['b3']


@pytest.mark.parametrize('d_in,n_out,act,rows', [
    (17, 1, 'tanh', 16384), (17, 6, 'tanh', 1000), (23, 1, 'relu', 333), (32, 8, 'relu', 128),
    (3, 1, 'tanh', 1), (17, 6, 'tanh', 148 * 128 * 2 + 77)])
@pytest.mark.parametrize('passes', [3, 1])
def test_tc_mlp_forward_fused(K, d_in, n_out, act, rows, passes):
    """tb_tc_mlp_forward (input gather + normalisation, both hidden layers and the head in one
    tcgen05 kernel) against a float64 evaluation of the same network."""
    import ctypes
    from tonic_b200 import _lib
    layout = K.MlpLayout(d_in, 256, n_out, act)
    net = K.DeviceMlp(layout)
    g = torch.Generator().manual_seed(rows + d_in)
    net.params.copy_(torch.randn(layout.n_params, generator=g) * 0.15)
    net.pack()
    pool = torch.randn(rows + 50, d_in, generator=g)
    idx = torch.randperm(rows + 50, generator=g)[:rows]
    mean, std = torch.randn(d_in, generator=g) * 0.1, torch.rand(d_in, generator=g) + 0.5
    x = (pool[idx] - mean) / std
    h1, h2, out = _reference_forward(net, x, act)
    inp = K.MlpInput(pool.cuda(), mean.cuda(), std.cuda(), idx=idx.cuda())
    tol = dict(rtol=3e-5, atol=3e-5) if passes == 3 else dict(rtol=2e-2, atol=2e-2)
    for save in (True, False):
        got = torch.full((rows, n_out), float('nan'), device='cuda')
        bufs = [torch.full((rows, 256), float('nan'), device='cuda') for _ in range(3)] if save else [None] * 3
        xin = torch.full((rows, layout.ldx), float('nan'), device='cuda') if save else None
        _lib.call('tb_tc_mlp_forward', ctypes.byref(layout.shape), K.ptr(net.params), K.ptr(net.packed),
                  ctypes.byref(inp.struct), rows, K.ptr(got), K.ptr(xin), K.ptr(bufs[0]),
                  K.ptr(bufs[1]), K.ptr(bufs[2]), passes, None, K.stream())
        torch.cuda.synchronize()
        np.testing.assert_allclose(got.cpu(), out, **tol)
        if save:
            h1_hi, h1_lo, h2_got = bufs
            # (1-pass mode stores only the tf32 part)
            np.testing.assert_allclose((h1_hi + h1_lo if passes == 3 else h1_hi).cpu(), h1, **tol)
            np.testing.assert_allclose(h2_got.cpu(), h2, **tol)
            assert (h1_hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0
            want = torch.cat([x, torch.ones(rows, 1), torch.zeros(rows, layout.ldx - d_in - 1)], 1)
            np.testing.assert_allclose(xin.cpu(), want, rtol=1e-6, atol=1e-6)


def test_tc_mlp_forward_two_inputs(K):
    """Q-critic style input: [observations | actions] with the second part not gathered."""
    import ctypes
    from tonic_b200 import _lib
    rows, d1, d2 = 500, 11, 4
    layout = K.MlpLayout(d1 + d2, 256, 1, 'relu')
    net = K.DeviceMlp(layout)
    g = torch.Generator().manual_seed(5)
    net.params.copy_(torch.randn(layout.n_params, generator=g) * 0.15)
    net.pack()
    pool = torch.randn(900, d1, generator=g)
    idx = torch.randint(900, (rows,), generator=g)
    acts = torch.randn(rows, d2, generator=g)
    _, _, out = _reference_forward(net, torch.cat([pool[idx], acts], 1), 'relu')
    inp = K.MlpInput(pool.cuda(), x2=acts.cuda(), gather2=False, idx=idx.cuda())
    got = torch.empty(rows, 1, device='cuda')
    _lib.call('tb_tc_mlp_forward', ctypes.byref(layout.shape), K.ptr(net.params), K.ptr(net.packed),
              ctypes.byref(inp.struct), rows, K.ptr(got), None, None, None, None, 3, None, K.stream())
    torch.cuda.synchronize()
    np.testing.assert_allclose(got.cpu(), out, rtol=3e-5, atol=3e-5)


@pytest.mark.parametrize('n_out,act,rows', [(1, 'tanh', 16384), (6, 'tanh', 1000), (3, 'relu', 77),
                                            (8, 'relu', 148 * 128 * 2 + 77)])
@pytest.mark.parametrize('passes', [3, 1])
def test_tc_mlp_backward_fused(K, n_out, act, rows, passes):
    """tb_tc_mlp_backward (head gradient, hidden-layer GEMM and both activation gradients in
    one tcgen05 kernel) against float64."""
    import ctypes
    from tonic_b200 import _lib
    layout = K.MlpLayout(17, 256, n_out, act)
    net = K.DeviceMlp(layout)
    g = torch.Generator().manual_seed(rows + n_out)
    net.params.copy_(torch.randn(layout.n_params, generator=g) * 0.15)
    net.pack()
    f = torch.tanh if act == 'tanh' else torch.relu
    h1 = f(torch.randn(rows, 256, generator=g))
    h2 = f(torch.randn(rows, 256, generator=g))
    dout = torch.randn(rows, n_out + 2, generator=g)          # wider rows: ld_dout > n_out
    w2 = net.view('w2', (256, 256)).cpu().double()
    w3 = net.view('w3', (n_out, 256)).cpu().double()
    grad = (lambda h: 1 - h * h) if act == 'tanh' else (lambda h: (h > 0).double())
    dz2 = (dout[:, :n_out].double() @ w3) * grad(h2.double())
    dz1 = (dz2 @ w2) * grad(h1.double())
    h1_hi, h1_lo = split(K, h1.cuda())
    out = [torch.full((rows, 256), float('nan'), device='cuda') for _ in range(3)]
    d_dout, d_h2 = dout.cuda(), h2.cuda()        # keep the device copies alive across the call
    _lib.call('tb_tc_mlp_backward', ctypes.byref(layout.shape), K.ptr(net.params), K.ptr(net.packed),
              K.ptr(d_dout), n_out + 2, K.ptr(h1_hi), K.ptr(h1_lo), K.ptr(d_h2), rows,
              K.ptr(out[0]), K.ptr(out[1]), K.ptr(out[2]), passes, None, K.stream())
    torch.cuda.synchronize()
    # 1-pass mode keeps only the tf32 part of dz2 (2^-11 truncation), 3-pass mode the exact split
    t2 = 1e-5 if passes == 3 else 1e-3
    np.testing.assert_allclose((out[0] + out[1] if passes == 3 else out[0]).cpu(), dz2, rtol=t2, atol=t2)
    assert (out[0].view(torch.int32) & 0x1FFF).abs().max().item() == 0
    tol = 1e-5 if passes == 3 else 5e-3
    err = (out[2].cpu().double() - dz1).abs().max().item()
    assert err <= tol * dz1.abs().max().item(), (err, dz1.abs().max().item())


@pytest.mark.parametrize('d_in,n_out,n_extra,rows,n_split', [
    (17, 1, 0, 16384, 74), (17, 6, 6, 16384, 74), (23, 8, 0, 1000, 31), (31, 3, 2, 77, 2),
    (3, 1, 0, 33, 1), (17, 6, 6, 148 * 128 + 5, 74)])
@pytest.mark.parametrize('passes', [3, 1])
def test_wgrad_fused_all_gradients_one_launch(K, d_in, n_out, n_extra, rows, n_split, passes):
    """tb_mlp_wgrad_fused (tcgen05 dW2 / db2 + FFMA narrow gradients in the same CTAs + grid
    barrier + fixed-order reduction to the flat gradient) against float64 sums; two launches
    give bit-identical results (deterministic reduction order)."""
    import ctypes
    from tonic_b200 import _lib
    layout = K.MlpLayout(d_in, 256, n_out, 'tanh', [('log_scale', n_extra)] if n_extra else ())
    sh = layout.shape
    g = torch.Generator().manual_seed(rows + d_in)
    xin = torch.zeros(rows, layout.ldx)
    xin[:, :d_in] = torch.randn(rows, d_in, generator=g)
    xin[:, d_in] = 1.0
    h1, h2 = torch.tanh(torch.randn(rows, 256, generator=g)), torch.tanh(torch.randn(rows, 256, generator=g))
    dz1, dz2 = torch.randn(rows, 256, generator=g) * 0.1, torch.randn(rows, 256, generator=g) * 0.1
    ld = n_out + n_extra + 1
    dout = torch.randn(rows, ld, generator=g)
    ref = torch.zeros(layout.n_params, dtype=torch.float64)
    o = layout.offsets
    d = lambda t: t.double()      # noqa: E731
    ref[o['w1'][0]:o['w1'][0] + 256 * d_in] = (d(dz1).T @ d(xin[:, :d_in])).reshape(-1)
    ref[o['b1'][0]:o['b1'][0] + 256] = d(dz1).sum(0)
    ref[o['w2'][0]:o['w2'][0] + 65536] = (d(dz2).T @ d(h1)).reshape(-1)
    ref[o['b2'][0]:o['b2'][0] + 256] = d(dz2).sum(0)
    ref[o['w3'][0]:o['w3'][0] + n_out * 256] = (d(dout[:, :n_out]).T @ d(h2)).reshape(-1)
    ref[o['b3'][0]:o['b3'][0] + n_out] = d(dout[:, :n_out]).sum(0)
    off_extra = 0
    if n_extra:
        off_extra = o['log_scale'][0]
        ref[off_extra:off_extra + n_extra] = d(dout[:, n_out:n_out + n_extra]).sum(0)
    dev = [t.cuda() for t in (xin, h2, dz1, dout)]
    h1_hi, h1_lo = split(K, h1.cuda())
    dz2_hi, dz2_lo = split(K, dz2.cuda())
    gpart = torch.zeros(n_split, layout.n_params, device='cuda')
    sync = torch.zeros(1, dtype=torch.int64, device='cuda')
    outs = []
    for _ in range(2):
        flat = torch.full((layout.n_params,), float('nan'), device='cuda')
        _lib.call('tb_mlp_wgrad_fused', ctypes.byref(sh), K.ptr(dev[0]), K.ptr(h1_hi), K.ptr(h1_lo),
                  K.ptr(dev[1]), K.ptr(dev[2]), K.ptr(dz2_hi), K.ptr(dz2_lo), K.ptr(dev[3]), ld, n_extra,
                  off_extra, rows, K.ptr(gpart), n_split, K.ptr(flat), K.ptr(sync), passes,
                  None, None, 0.0, None, -1.0, None, None, None, None, None, K.stream())
        torch.cuda.synchronize()
        outs.append(flat.cpu())
    assert torch.equal(outs[0], outs[1])
    assert int(sync.item()) == 2 << 32                  # generation 2, no arrivals pending
    if passes == 3:
        # plain float32 h1 / dz2 (lo == NULL): the kernel splits the tiles in shared memory into
        # the same hi / lo operands -> bit-identical gradient
        d_h1, d_dz2 = h1.cuda(), dz2.cuda()
        flat = torch.full((layout.n_params,), float('nan'), device='cuda')
        _lib.call('tb_mlp_wgrad_fused', ctypes.byref(sh), K.ptr(dev[0]), K.ptr(d_h1), None,
                  K.ptr(dev[1]), K.ptr(dev[2]), K.ptr(d_dz2), None, K.ptr(dev[3]), ld, n_extra,
                  off_extra, rows, K.ptr(gpart), n_split, K.ptr(flat), K.ptr(sync), passes,
                  None, None, 0.0, None, -1.0, None, None, None, None, None, K.stream())
        torch.cuda.synchronize()
        assert torch.equal(flat.cpu(), outs[0])
    got = outs[0].double()
    used = torch.zeros(layout.n_params, dtype=torch.bool)
    for name, (off, size) in o.items():
        used[off:off + size] = True
    scale = ref.abs().max().item()
    tol = 2e-6 * scale * (rows ** 0.5) if passes == 3 else 2e-3 * scale
    narrow = torch.ones(layout.n_params, dtype=torch.bool)
    narrow[o['w2'][0]:o['w2'][0] + 65536 + 256] = False
    err_n = (got - ref)[used & narrow].abs().max().item()
    err_w = (got - ref)[used & ~narrow].abs().max().item()
    assert err_n <= 2e-6 * scale * (rows ** 0.5) + 1e-5, (err_n, scale)       # FFMA part: fp32 in any mode
    assert err_w <= tol + 1e-5, (err_w, scale)


@pytest.mark.timeout(240, method='thread')
def test_wgrad_fused_grid_barrier_survives_changing_grid_sizes(K):
    """Consecutive launches on ONE barrier word with different numbers of CTAs (a ragged last
    minibatch changes the split count from launch to launch): every launch must complete and give
    the gradient of its own rows.  (A counter that only grows deadlocks here: the arrival count is
    not a multiple of the new grid size.)"""
    import ctypes
    from tonic_b200 import _lib
    layout = K.MlpLayout(17, 256, 1, 'tanh')
    g = torch.Generator().manual_seed(11)
    big = 2048
    xin = torch.zeros(big, layout.ldx)
    xin[:, :17] = torch.randn(big, 17, generator=g)
    xin[:, 17] = 1.0
    h1, h2 = torch.tanh(torch.randn(big, 256, generator=g)).cuda(), torch.tanh(torch.randn(big, 256, generator=g)).cuda()
    dz1, dz2 = (torch.randn(big, 256, generator=g) * 0.1).cuda(), (torch.randn(big, 256, generator=g) * 0.1).cuda()
    dout = torch.randn(big, 2, generator=g).cuda()
    xin = xin.cuda()
    sync = torch.zeros(1, dtype=torch.int64, device='cuda')
    gpart = torch.zeros(74, layout.n_params, device='cuda')
    o = layout.offsets
    launches = [(64, 2), (32, 1), (64, 2), (2048, 64), (33, 1), (1000, 31), (64, 2), (2048, 74), (48, 1)]
    for k, (rows, n_split) in enumerate(launches):
        flat = torch.full((layout.n_params,), float('nan'), device='cuda')
        _lib.call('tb_mlp_wgrad_fused', ctypes.byref(layout.shape), K.ptr(xin), K.ptr(h1), None,
                  K.ptr(h2), K.ptr(dz1), K.ptr(dz2), None, K.ptr(dout), 2, 0, 0, rows, K.ptr(gpart), n_split,
                  K.ptr(flat), K.ptr(sync), 3, None, None, 0.0, None, -1.0, None, None, None, None, None,
                  K.stream())
        torch.cuda.synchronize()
        assert int(sync.item()) == (k + 1) << 32
        want = (dz2[:rows].double().T @ h1[:rows].double()).reshape(-1)
        got = flat[o['w2'][0]:o['w2'][0] + 65536].double()
        assert (got - want).abs().max().item() <= 2e-6 * want.abs().max().item() * rows ** 0.5 + 1e-5, (rows, n_split)
        want_b1 = dz1[:rows].double().sum(0)
        got_b1 = flat[o['b1'][0]:o['b1'][0] + 256].double()
        assert (got_b1 - want_b1).abs().max().item() <= 1e-4, (rows, n_split)


@pytest.mark.parametrize('n_out,n_extra,with_stats', [(1, 0, False), (6, 6, True)])
def test_wgrad_fused_adam_equals_separate_step(K, n_out, n_extra, with_stats):
    """The optimizer step inside tb_mlp_wgrad_fused (reduction phase) is bit-identical to
    tb_adam_step on the flat gradient: parameters, moments, packed operands, step counter and
    the device-side PPO controls (zero-advantage skip, KL stop flag)."""
    import ctypes
    from tonic_b200 import _lib
    rows, n_split, d_in = 4096, 74, 17
    extras = [('log_scale', n_extra)] if n_extra else ()
    g = torch.Generator().manual_seed(7)
    data = dict(xin=torch.randn(rows, 20, generator=g), h2=torch.randn(rows, 256, generator=g),
                dz1=torch.randn(rows, 256, generator=g) * 0.1, dout=torch.randn(rows, n_out + n_extra, generator=g),
                h1=torch.tanh(torch.randn(rows, 256, generator=g)), dz2=torch.randn(rows, 256, generator=g) * 0.1)
    dev = {k: v.cuda() for k, v in data.items()}
    h1_hi, h1_lo = split(K, dev['h1'])
    dz2_hi, dz2_lo = split(K, dev['dz2'])
    init = torch.randn(K.MlpLayout(d_in, 256, n_out, 'tanh', extras).n_params, generator=g) * 0.1

    def run(fused, stats_vec):
        layout = K.MlpLayout(d_in, 256, n_out, 'tanh', extras)
        net = K.DeviceMlp(layout)
        net.params.copy_(init)
        net.pack()
        adam = K.Adam(net.params, lr=1e-3)
        adam.m.copy_(init * 0.01)
        adam.v.copy_(init.abs() * 0.001)
        gpart = torch.zeros(n_split, layout.n_params, device='cuda')
        flat = torch.zeros(layout.n_params, device='cuda')
        sync = torch.zeros(1, dtype=torch.int64, device='cuda')
        stats = None if stats_vec is None else torch.tensor(stats_vec, dtype=torch.float64, device='cuda')
        stop = torch.zeros(1, dtype=torch.int32, device='cuda')
        off_extra = layout.offsets['log_scale'][0] if n_extra else 0
        for _ in range(2):        # two steps: the step counter advances
            opt = ctypes.byref(adam.struct) if fused else None
            _lib.call('tb_mlp_wgrad_fused', ctypes.byref(layout.shape), K.ptr(dev['xin']), K.ptr(h1_hi),
                      K.ptr(h1_lo), K.ptr(dev['h2']), K.ptr(dev['dz1']), K.ptr(dz2_hi), K.ptr(dz2_lo),
                      K.ptr(dev['dout']), n_out + n_extra, n_extra, off_extra, rows, K.ptr(gpart), n_split,
                      K.ptr(flat), K.ptr(sync), 3, opt, K.ptr(net.packed) if fused else None,
                      1.0 / rows if fused else 0.0, K.ptr(stats) if fused else None,
                      0.015 if fused and stats is not None else -1.0, K.ptr(stop) if fused else None,
                      None, None, None, None, K.stream())
            if not fused:
                adam.step(net, flat, 1, 1.0 / rows, stats=stats, kl_threshold=0.015 if stats is not None else -1.0,
                          stop=stop)
        torch.cuda.synchronize()
        return [t.cpu().clone() for t in (net.params, adam.m, adam.v, net.packed, adam.step_count, stop)]

    cases = [None]
    if with_stats:
        live = [0.0] * _lib.STAT_COUNT
        live[_lib.STAT_ROWS], live[_lib.STAT_NONZERO_ADV], live[_lib.STAT_KL] = rows, rows, 0.02 * rows
        dead = list(live)
        dead[_lib.STAT_NONZERO_ADV] = 0.0
        cases = [live, dead]
    for stats_vec in cases:
        a, b = run(True, stats_vec), run(False, stats_vec)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        if stats_vec is None or stats_vec[_lib.STAT_NONZERO_ADV] > 0:
            assert int(a[4][0]) == 2 and not torch.equal(a[0], init)
            assert int(a[5][0]) == (1 if stats_vec is not None else 0)       # kl 0.02 > 0.015 -> stop
        else:
            assert int(a[4][0]) == 0 and torch.equal(a[0], init) and int(a[5][0]) == 0


@pytest.mark.parametrize('kind,n_out,act,rows', [
    ('value', 1, 'tanh', 16384), ('policy', 6, 'tanh', 16384), ('value', 1, 'relu', 1000),
    ('policy', 3, 'tanh', 77), ('policy', 8, 'relu', 148 * 128 + 9), ('a2c', 6, 'tanh', 333)])
def test_tc_mlp_train_equals_forward_loss_backward(K, kind, n_out, act, rows):
    """tb_tc_mlp_train (forward -> loss -> backward in one launch, z2 kept in TMEM) against the
    three-kernel chain tb_tc_mlp_forward -> tb_gauss_policy_loss / fused MSE -> tb_tc_mlp_backward:
    same arithmetic -> bit-identical activations and gradients, statistics to double round-off."""
    from tonic_b200 import _lib
    d_in = 17
    policy = kind != 'value'
    extras = [('log_scale', n_out)] if policy else ()
    g = torch.Generator().manual_seed(rows + n_out)
    layout = K.MlpLayout(d_in, 256, n_out, act, extras)
    init = torch.randn(layout.n_params, generator=g) * 0.12
    pool = torch.randn(rows + 40, d_in, generator=g).cuda()
    idx = torch.randperm(rows + 40, generator=g)[:rows].cuda()
    mean, std = (torch.randn(d_in, generator=g) * 0.1).cuda(), (torch.rand(d_in, generator=g) + 0.5).cuda()
    targets = torch.randn(rows + 40, generator=g).cuda()
    actions = torch.randn(rows + 40, n_out, generator=g).cuda()
    adv = torch.randn(rows + 40, generator=g).cuda()
    adv[idx[:5]] = 0.0
    old_lp = (torch.randn(rows + 40, generator=g) * 0.3 - 5.0).cuda()
    ratio_clip = 0.0 if kind == 'a2c' else 0.2
    results = []
    for fused in (False, True):
        net = K.DeviceMlp(layout)
        net.params.copy_(init)
        net.pack()
        stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.float64, device='cuda')
        dout = torch.zeros(rows, 2 * n_out if policy else 1, device='cuda')
        out = torch.zeros(rows, n_out, device='cuda')
        inp = K.MlpInput(pool, mean, std, idx=idx)
        if fused:
            assert net.fused_train()
            if policy:
                off, size = layout.offsets['log_scale']
                net.train_step(inp, rows, dout, stats, idx=idx, out=out,
                               policy=dict(log_scale=net.params[off:off + size], actions=actions, advantages=adv,
                                           log_probs=old_lp, ratio_clip=ratio_clip, entropy_coeff=0.01))
            else:
                net.train_step(inp, rows, dout, stats, idx=idx, targets=targets, out=out)
        else:
            if policy:
                off, size = layout.offsets['log_scale']
                net.forward(inp, rows, out, save=True)
                K.gauss_policy_loss(out, net.params[off:off + size], actions, adv, old_lp, idx, rows, dout,
                                    stats, ratio_clip, 0.01)
            else:
                net.forward(inp, rows, out, save=True, vloss=(targets, idx, dout, stats))
                assert net.vloss_fused
            net.backward(dout, rows)
        torch.cuda.synchronize()
        acts = [net.xin[:rows], net.h1[:rows], net.h2[:rows], net.dz2[:rows], net.dz1[:rows]]
        if not net.plain_activations():
            acts += [net.h1_lo[:rows], net.dz2_lo[:rows]]
        results.append(([t.clone() for t in acts], dout.clone(), out.clone(), stats.cpu().numpy()))
    (a0, d0, o0, s0), (a1, d1, o1, s1) = results
    assert torch.equal(o0, o1) and torch.equal(d0, d1)
    for x, y in zip(a0, a1):
        assert torch.equal(x, y)
    assert float(a0[4].abs().max()) > 0 and torch.isfinite(a0[4]).all()
    np.testing.assert_allclose(s1, s0, rtol=1e-12, atol=1e-12)
    assert s0[_lib.STAT_ROWS] == rows
