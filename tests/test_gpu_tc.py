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
    a = torch.randn(rows, 256) * 0.3
    w = torch.randn(256, 256) * 0.1
    bias = torch.randn(256) * 0.1
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
    np.testing.assert_allclose((out_hi + out_lo).cpu(), ref, rtol=2e-5, atol=2e-5)
    assert (out_hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0
    # backward: dz1 = (dz2 W) * act'(h1): B operand is W^T (rows = output index)
    h1 = f(torch.randn(rows, 256))
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
    n_params, off = 70000, 4352
    gpart = torch.full((n_split, n_params), float('nan'), device='cuda')
    K.tc_wgrad256(dz_hi, dz_lo, h_hi, h_lo, rows, gpart, n_split, n_params, off, passes=passes)
    got = gpart[:, off:off + 65536].sum(0).view(256, 256).cpu().double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= (1e-5 if passes == 3 else 5e-3) * scale, (err, scale)
    # nothing outside the W2 block is touched
    assert torch.isnan(gpart[:, :off]).all() and torch.isnan(gpart[:, off + 65536:]).all()


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
