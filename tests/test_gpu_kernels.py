"""Kernel-level parity: each sm_100a kernel, called through the C ABI, against the
oracle port / golden vectors / a plain fp32 torch-CPU restatement of the op.
Integer and exactly-rounded paths are compared bit-for-bit; floating-point
paths with the tolerance stated at each assert."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import port, synth_env  # noqa: E402  (checker only)


@pytest.fixture(scope='module')
def K():
    from tonic_b200 import kernels
    kernels.device()
    return kernels


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x)).to('cuda', dtype)


# ------------------------------------------------------------------ environment
def make_env(obs, act, workers, max_steps, seed, first_worker=0):
    from tonic_b200 import environments
    spec = environments.SynthControl('synth', obs, act, max_steps)
    env = environments.DeviceVectorEnvironment(spec, workers, first_worker=first_worker)
    env.initialize(seed)
    return env


@pytest.mark.parametrize('obs,act,workers,max_steps', [(17, 6, 37, 7), (376, 17, 9, 5), (3, 1, 300, 4)])
def test_env_bit_exact_vs_oracle(K, obs, act, workers, max_steps):
    env = make_env(obs, act, workers, max_steps, seed=5)
    ref = port.VectorEnv(obs, act, workers, max_steps)
    ref.initialize(5)
    o_dev, o_ref = env.start(host=True), ref.start()
    np.testing.assert_array_equal(o_dev, o_ref)
    rs = np.random.RandomState(0)
    scores = np.zeros(workers)
    lengths = np.zeros(workers, int)
    fin_scores, fin_lengths = [], []
    for t in range(25):
        a = (rs.normal(size=(workers, act)) * 1.5).astype(np.float32)
        o_dev, i_dev = env.step(a)
        o_ref, i_ref = ref.step(a)
        np.testing.assert_array_equal(o_dev, o_ref)
        for k in ('observations', 'rewards', 'resets', 'terminations'):
            np.testing.assert_array_equal(i_dev[k], i_ref[k], err_msg=f'{k} step {t}')
        scores += i_ref['rewards']          # trainer.py:52-71 bookkeeping
        lengths += 1
        for i in range(workers):
            if i_ref['resets'][i]:
                fin_scores.append(scores[i]); fin_lengths.append(lengths[i])
                scores[i] = 0; lengths[i] = 0
    got_s, got_l = env.finished_episodes()
    assert len(got_s) == len(fin_scores) > 0
    np.testing.assert_allclose(np.sort(got_s), np.sort(fin_scores), rtol=1e-12)
    np.testing.assert_array_equal(np.sort(got_l), np.sort(fin_lengths))


def test_env_golden_reference_sequential(K, golden):
    g = golden('units')
    env = make_env(7, 3, 5, 6, seed=21)
    obs = [env.start(host=True)]
    for t, a in enumerate(g['env/actions']):
        o, infos = env.step(a)
        obs.append(o)
        np.testing.assert_array_equal(infos['observations'], g['env/next_observations'][t])
        np.testing.assert_array_equal(infos['rewards'], g['env/rewards'][t])
        np.testing.assert_array_equal(infos['resets'], g['env/resets'][t])
        np.testing.assert_array_equal(infos['terminations'], g['env/terminations'][t])
    np.testing.assert_array_equal(np.array(obs), g['env/observations'])


def test_env_sharding_matches_single(K):
    whole = make_env(17, 6, 12, 9, seed=3)
    parts = [make_env(17, 6, 6, 9, seed=3, first_worker=f) for f in (0, 6)]
    o = whole.start(host=True)
    op = np.concatenate([p.start(host=True) for p in parts])
    np.testing.assert_array_equal(o, op)
    rs = np.random.RandomState(1)
    for _ in range(12):
        a = rs.normal(size=(12, 6)).astype(np.float32)
        o, i = whole.step(a)
        outs = [p.step(a[k * 6:(k + 1) * 6]) for k, p in enumerate(parts)]
        np.testing.assert_array_equal(o, np.concatenate([x[0] for x in outs]))
        np.testing.assert_array_equal(i['rewards'], np.concatenate([x[1]['rewards'] for x in outs]))


# ------------------------------------------------------------------ returns
def test_lambda_returns_golden_bit_exact(K, golden):
    g = golden('units')
    for tag in 'abc':
        args = [dev(g[f'lam_{tag}/{k}']) for k in
                ('values', 'next_values', 'rewards', 'resets', 'terminations')]
        out = torch.empty_like(args[0])
        K.lambda_returns(*args, out, 0.99, 0.97)
        np.testing.assert_array_equal(out.cpu().numpy(), g[f'lam_{tag}/returns'])


def test_lambda_returns_large_bit_exact(K):
    rs = np.random.RandomState(4)
    T, N = 128, 1000
    v, nv, r = (rs.normal(size=(T, N)).astype(np.float32) for _ in range(3))
    resets = (rs.uniform(size=(T, N)) < 0.05).astype(np.float32)
    terms = (resets * (rs.uniform(size=(T, N)) < 0.5)).astype(np.float32)
    out = torch.empty(T, N, device='cuda')
    K.lambda_returns(dev(v), dev(nv), dev(r), dev(resets), dev(terms), out, 0.99, 0.97)
    np.testing.assert_array_equal(out.cpu().numpy(),
                                  port.lambda_returns(v, nv, r, resets, terms, 0.99, 0.97))


def test_advantages(K, golden):
    g = golden('units')
    ret, val = g['adv/returns'], g['adv/values'].reshape(g['adv/returns'].shape)
    out = torch.empty(ret.size, device='cuda')
    ws = torch.zeros(4, dtype=torch.float64, device='cuda')
    K.advantages(dev(ret.ravel()), dev(val.ravel()), out, ws)
    # float32 elementwise after fp64 statistics: 1e-5 relative (numpy uses f32 pairwise sums)
    np.testing.assert_allclose(out.cpu().numpy(), g['adv/advantages'], rtol=1e-5, atol=1e-6)
    # std == 0 -> advantages left un-normalised (segments.py:44)
    const = np.full(64, 2.5, np.float32)
    K.advantages(dev(const), dev(const * 0 + 1), out[:64], ws)
    np.testing.assert_array_equal(out[:64].cpu().numpy(), np.full(64, 1.5, np.float32))


def test_moments(K, golden):
    g = golden('units')
    dim = 5
    sums = torch.zeros(2 * dim + 1, dtype=torch.float64, device='cuda')
    running = torch.zeros(2 * dim, device='cuda')
    count = torch.zeros(1, dtype=torch.float64, device='cuda')
    mean, std = torch.zeros(dim, device='cuda'), torch.ones(dim, device='cuda')
    snaps = []
    for i, b in enumerate(g['meanstd/batches']):
        K.moments_record(dev(b), sums)
        if i % 2 == 1:
            K.moments_update(sums, running, count, mean, std)
            snaps.append(np.stack([mean.cpu().numpy(), std.cpu().numpy()]))
    # reference sums in float32 sequentially, the kernel in float64: 1e-5 relative
    np.testing.assert_allclose(np.stack(snaps), g['meanstd/snapshots'], rtol=1e-5, atol=1e-6)
    assert float(count.item()) == 42
    # wide rows (dim > 256) and the std floor
    x = np.random.RandomState(0).normal(size=(700, 376)).astype(np.float32)
    x[:, 7] = 3.0
    sums = torch.zeros(2 * 376 + 1, dtype=torch.float64, device='cuda')
    running = torch.zeros(2 * 376, device='cuda')
    mean, std = torch.zeros(376, device='cuda'), torch.ones(376, device='cuda')
    K.moments_record(dev(x), sums)
    K.moments_update(sums, running, count.zero_(), mean, std)
    np.testing.assert_allclose(mean.cpu().numpy(), x.mean(0), rtol=1e-4, atol=1e-5)
    ref_std = np.maximum(x.std(0), 1e-2)
    np.testing.assert_allclose(std.cpu().numpy(), ref_std, rtol=1e-3, atol=1e-4)
    assert std[7].item() == pytest.approx(1e-2)


# ------------------------------------------------------------------ MLP
@pytest.fixture(params=['ffma', 'tf32x3'])
def gemm_mode(request):
    """Both GEMM paths: FP32 FFMA kernels and the tcgen05 3xTF32 tensor-core path (the
    latter only changes the 256-wide configurations)."""
    from tonic_b200 import config
    old = config.gemm
    config.gemm = request.param
    yield request.param
    config.gemm = old


def make_mlp(K, d_in, hidden, n_out, act, extras=(), seed=0):
    layout = K.MlpLayout(d_in, hidden, n_out, act, extras)
    net = K.DeviceMlp(layout)
    g = torch.Generator().manual_seed(seed)
    net.params.copy_(torch.randn(layout.n_params, generator=g) * 0.2)
    net.pack()
    return net


def host_params(net):
    L = net.layout
    H = L.hidden
    p = {k: net.view(k, s).cpu().clone() for k, s in
         dict(w1=(H, L.d_in), b1=(H,), w2=(H, H), b2=(H,), w3=(L.n_out, H), b3=(L.n_out,)).items()}
    return p


def torch_forward(p, x, act):
    f = torch.tanh if act == 'tanh' else torch.relu
    h1 = f(x @ p['w1'].T + p['b1'])
    h2 = f(h1 @ p['w2'].T + p['b2'])
    return h1, h2, h2 @ p['w3'].T + p['b3']


@pytest.mark.parametrize('d_in,hidden,n_out,act,rows', [
    (17, 64, 6, 'tanh', 1), (17, 256, 6, 'tanh', 200), (17, 256, 1, 'tanh', 64),
    (5, 128, 2, 'relu', 63), (393, 256, 1, 'relu', 130), (28, 64, 34, 'relu', 65),
    (393, 256, 1, 'relu', 4200), (40, 256, 34, 'tanh', 90), (40, 256, 34, 'relu', 4200)])
def test_mlp_forward(K, gemm_mode, d_in, hidden, n_out, act, rows):
    net = make_mlp(K, d_in, hidden, n_out, act)
    x = torch.randn(rows, d_in)
    out = torch.empty(rows, n_out, device='cuda')
    net.forward(K.MlpInput(x.cuda()), rows, out, save=True)
    h1, h2, ref = torch_forward(host_params(net), x, act)
    # fp32 FFMA vs fp32 CPU GEMM: 2e-5 relative to the row scale
    tol = dict(rtol=2e-5, atol=2e-5 * float(ref.abs().max()))
    np.testing.assert_allclose(out.cpu(), ref, **tol)
    split = net.passes() and not net.plain_activations()      # tf32 hi / lo pair or one float32 array
    got_h1 = net.h1[:rows] + net.h1_lo[:rows] if split else net.h1[:rows]
    # hidden activations: errors scale with the magnitude of the pre-activation row
    np.testing.assert_allclose(got_h1.cpu(), h1, rtol=2e-5, atol=2e-5 * max(1.0, float(h1.abs().max())))
    np.testing.assert_allclose(net.h2[:rows].cpu(), h2, rtol=2e-5,
                               atol=2e-5 * max(1.0, float(h2.abs().max())))
    np.testing.assert_array_equal(net.xin[:rows, :d_in].cpu(), x)
    np.testing.assert_array_equal(net.xin[:rows, d_in].cpu(), torch.ones(rows))


@pytest.mark.parametrize('d_in,n_out,act,rows', [(393, 1, 'relu', 100), (376, 34, 'relu', 100),
                                                  (111, 8, 'tanh', 257), (40, 34, 'tanh', 3)])
def test_small_batch_kernels_equal_the_tile_kernels(K, d_in, n_out, act, rows):
    """Off-policy minibatches (100 rows, replays/buffers.py:8-12) run the wide first layer, the
    input gradient and wide heads on small-CTA kernels; they must reproduce the 64-row tile kernels
    bit for bit (same fmaf chain per output), so that results do not depend on the batch size."""
    from tonic_b200 import _lib, config
    if config.gemm == 'ffma':
        pytest.skip('tensor-core chain only')
    got = {}
    g = torch.Generator().manual_seed(d_in + rows)
    x = torch.randn(rows + 20, d_in - 5, generator=g).cuda()
    x2 = torch.randn(rows, 5, generator=g).cuda()
    idx = torch.randperm(rows + 20, generator=g)[:rows].cuda()
    mean, std = torch.randn(d_in - 5, generator=g).cuda(), (torch.rand(d_in - 5, generator=g) + 0.5).cuda()
    dout = torch.randn(rows, K.round_up(n_out, 4), generator=g).cuda()
    try:
        for on in (1, 0):
            _lib.call('tb_debug_skinny', on)
            net = make_mlp(K, d_in, 256, n_out, act)
            out = torch.empty(rows, n_out, device='cuda')
            net.forward(K.MlpInput(x, mean, std, x2=x2, gather2=False, idx=idx), rows, out, save=True)
            dx = torch.empty(rows, 5, device='cuda')
            net.backward(dout, rows, dx=dx, dx_col0=d_in - 5)
            torch.cuda.synchronize()
            got[on] = [t[:rows].clone() for t in (out, net.h1, net.h1_lo, net.h2, net.xin, dx)]
    finally:
        _lib.call('tb_debug_skinny', 1)
    for a, b, name in zip(got[1], got[0], ('out', 'h1_hi', 'h1_lo', 'h2', 'xin', 'dx')):
        assert torch.isfinite(a).all(), name
        assert torch.equal(a, b), name


def test_mlp_forward_gather_normalise_concat(K, gemm_mode):
    obs_dim, act_dim, rows, pool = 11, 3, 100, 400
    net = make_mlp(K, obs_dim + act_dim, 256, 1, 'relu')
    obs, acts = torch.randn(pool, obs_dim) * 3 + 1, torch.randn(pool, act_dim)
    mean, std = torch.randn(obs_dim), torch.rand(obs_dim) + 0.5
    idx = torch.randint(0, pool, (rows,))
    out = torch.empty(rows, 1, device='cuda')
    inp = K.MlpInput(obs.cuda(), mean.cuda(), std.cuda(), x2=acts.cuda(), gather2=True,
                     idx=idx.cuda())
    net.forward(inp, rows, out)
    x = torch.cat([(obs[idx] - mean) / std, acts[idx]], -1)
    ref = torch_forward(host_params(net), x, 'relu')[2]
    np.testing.assert_allclose(out.cpu(), ref, rtol=2e-5, atol=2e-5 * float(ref.abs().max()))
    # second source not gathered (e.g. freshly computed actions, minibatch order)
    fresh = torch.randn(rows, act_dim)
    inp = K.MlpInput(obs.cuda(), mean.cuda(), std.cuda(), x2=fresh.cuda(), gather2=False,
                     idx=idx.cuda())
    net.forward(inp, rows, out)
    x = torch.cat([(obs[idx] - mean) / std, fresh], -1)
    ref = torch_forward(host_params(net), x, 'relu')[2]
    np.testing.assert_allclose(out.cpu(), ref, rtol=2e-5, atol=2e-5 * float(ref.abs().max()))


@pytest.mark.parametrize('d_in,hidden,n_out,act,rows,n_split', [
    (17, 64, 6, 'tanh', 100, 3), (17, 256, 6, 'tanh', 1000, 7), (17, 256, 1, 'tanh', 64, 1),
    (14, 256, 1, 'relu', 257, 4), (393, 256, 1, 'relu', 130, 2), (40, 128, 34, 'relu', 90, 5),
    (393, 256, 1, 'tanh', 4200, 3)])     # (tanh: no ReLU kinks to flip between fp32 evaluations)
def test_mlp_backward_wgrad_vs_autograd(K, gemm_mode, d_in, hidden, n_out, act, rows, n_split):
    n_extra = 3
    net = make_mlp(K, d_in, hidden, n_out, act, extras=[('extra', n_extra)])
    x = torch.randn(rows, d_in)
    out = torch.empty(rows, n_out, device='cuda')
    net.forward(K.MlpInput(x.cuda()), rows, out, save=True)
    ld = K.round_up(n_out + n_extra, 4)
    dout = torch.randn(rows, ld)
    dx_col0, dx_cols = max(0, d_in - 5), min(5, d_in)
    dx = torch.empty(rows, dx_cols, device='cuda')
    dout_d = dout.cuda()
    net.backward(dout_d, rows, dx=dx, dx_col0=dx_col0)
    off_extra = net.layout.offsets['extra'][0]
    gpart = net.wgrad(dout_d, rows, n_split, n_extra=n_extra, off_extra=off_extra)
    if getattr(net, 'reduced', False):      # fused kernel: already the flat gradient
        grad = gpart.clone()
    else:
        grad = gpart[:n_split].sum(0)
        w2_splits = net.w2_splits(n_split)
        if w2_splits:      # tensor-core path: the W2 block only has w2_splits partial sums
            lo, hi = net.w2_range()
            grad[lo:hi] = gpart[:w2_splits, lo:hi].sum(0)
    grad = grad.cpu()

    p = {k: v.requires_grad_() for k, v in host_params(net).items()}
    xr = x.clone().requires_grad_()
    ref_out = torch_forward(p, xr, act)[2]
    (ref_out * dout[:, :n_out]).sum().backward()
    scale = max(1.0, float(rows) ** 0.5)
    for k, shape in dict(w1=(hidden, d_in), b1=(hidden,), w2=(hidden, hidden), b2=(hidden,),
                         w3=(n_out, hidden), b3=(n_out,)).items():
        off, size = net.layout.offsets[k]
        got = grad[off:off + size].view(*shape)
        ref = p[k].grad
        # sums over `rows` fp32 products: tolerance scales with sqrt(rows)
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=2e-5 * scale * float(ref.abs().max() + 1),
                                   err_msg=k)
    np.testing.assert_allclose(grad[off_extra:off_extra + n_extra],
                               dout[:, n_out:n_out + n_extra].sum(0), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dx.cpu(), xr.grad[:, dx_col0:dx_col0 + dx_cols], rtol=1e-4,
                               atol=1e-5 * float(xr.grad.abs().max() + 1))


def test_adam_matches_torch_and_refreshes_packed(K):
    net = make_mlp(K, 17, 64, 6, 'tanh', extras=[('log_scale', 6)])
    P = net.layout.n_params
    ref_p = net.params.cpu().clone().requires_grad_()
    opt_ref = torch.optim.Adam([ref_p], lr=3e-4)
    opt = K.Adam(net.params, lr=3e-4)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        parts = torch.randn(3, P, generator=g)
        gpart = net.gpart(3)
        gpart.copy_(parts)
        opt.step(net, gpart, 3, 0.25)
        ref_p.grad = parts.sum(0) * 0.25
        opt_ref.step()
        # fp32 elementwise update, different summation order of 3 partials: 2e-6 abs
        np.testing.assert_allclose(net.params.cpu(), ref_p.detach(), rtol=1e-5, atol=2e-6)
    assert opt.step_count.cpu().tolist() == [4, 0]
    H = 64
    w1 = net.view('w1', (H, 17)).cpu()
    w2 = net.view('w2', (H, H)).cpu()
    np.testing.assert_array_equal(net.packed[:17 * H].view(17, H).cpu(), w1.T)
    np.testing.assert_array_equal(
        net.packed[net.layout.off_w2t:net.layout.off_w2t + H * H].view(H, H).cpu(), w2.T)
    # device-side skip flag
    before = net.params.clone()
    flag = torch.ones(1, dtype=torch.int32, device='cuda')
    opt.step(net, net.gpart(3), 3, 0.25, skip=flag)
    assert torch.equal(before, net.params) and opt.step_count[0].item() == 4


def test_soft_update(K):
    t, o = torch.randn(1000), torch.randn(1000)
    td = t.cuda()
    K.soft_update(td, o.cuda(), 0.005)
    ref = t.clone()
    ref.mul_(1 - 0.005)
    ref.add_(0.005 * o)
    np.testing.assert_array_equal(td.cpu(), ref)


# ------------------------------------------------------------------ heads
def ref_scale(log_scale):
    return torch.clamp(torch.nn.functional.softplus(log_scale) + 1e-8, 1e-4, 1.)


def test_gauss_sample_host_noise_and_philox(K):
    rows, A = 333, 6
    pre, ls, eps = torch.randn(rows, A), torch.randn(A) * 0.5, torch.randn(rows, A)
    actions = torch.empty(rows, A, device='cuda')
    logp = torch.empty(rows, device='cuda')
    K.gauss_sample(pre.cuda(), ls.cuda(), actions, logp, eps=eps.cuda())
    loc, scale = torch.tanh(pre), ref_scale(ls).expand(rows, A)
    dist = torch.distributions.Normal(loc, scale)
    ref_a = eps * scale + loc
    np.testing.assert_allclose(actions.cpu(), ref_a, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(logp.cpu(), dist.log_prob(actions.cpu()).sum(-1), rtol=1e-5, atol=1e-5)
    # device Philox noise: standardised samples ~ N(0, 1), reproducible per (seed, counter)
    rows = 20000
    pre = torch.zeros(rows, A, device='cuda')
    actions = torch.empty(rows, A, device='cuda')
    logp = torch.empty(rows, device='cuda')
    K.gauss_sample(pre, ls.cuda(), actions, logp, seed=7, counter=123)
    z = (actions.cpu() / ref_scale(ls)).numpy()
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.03
    again = torch.empty_like(actions)
    K.gauss_sample(pre, ls.cuda(), again, logp, seed=7, counter=123)
    assert torch.equal(actions, again)
    K.gauss_sample(pre, ls.cuda(), again, logp, seed=7, counter=123 + rows)
    assert not torch.equal(actions, again)


@pytest.mark.parametrize('ratio_clip,entropy_coeff', [(0.2, 0.0), (0.2, 0.01), (0.0, 0.0)])
def test_gauss_policy_loss_vs_autograd(K, ratio_clip, entropy_coeff):
    from tonic_b200 import _lib
    pool, rows, A = 500, 200, 6
    g = torch.Generator().manual_seed(3)
    pre = torch.randn(rows, A, generator=g).requires_grad_()
    ls = (torch.randn(A, generator=g) * 0.3).requires_grad_()
    actions = torch.randn(pool, A, generator=g)
    adv = torch.randn(pool, generator=g)
    adv[::7] = 0
    old = torch.randn(pool, generator=g) * 0.3 - 6
    idx = torch.randint(0, pool, (rows,), generator=g)
    dout = torch.empty(rows, 2 * A, device='cuda')
    stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.float64, device='cuda')
    K.gauss_policy_loss(pre.detach().cuda(), ls.detach().cuda(), actions.cuda(), adv.cuda(),
                        old.cuda(), idx.cuda(), rows, dout, stats, ratio_clip, entropy_coeff)
    # oracle formulas (updaters/actors.py:21-50,70-112), SUM loss
    dist = torch.distributions.Normal(torch.tanh(pre), ref_scale(ls).expand(rows, A))
    new = dist.log_prob(actions[idx]).sum(-1)
    if ratio_clip > 0:
        ratio = torch.exp(new - old[idx])
        lo, hi = 1 - ratio_clip, 1 + ratio_clip
        per_row = -torch.min(adv[idx] * ratio, adv[idx] * torch.clamp(ratio, lo, hi))
        clipped = (ratio.gt(hi) | ratio.lt(lo)).float().sum()
    else:
        per_row = -(adv[idx] * new)
        clipped = torch.tensor(0.)
    entropy = dist.entropy()
    loss = per_row.sum() - entropy_coeff * entropy.mean() * rows
    loss.backward()
    s = stats.cpu().numpy()
    np.testing.assert_allclose(dout[:, :A].cpu(), pre.grad, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(dout[:, A:].sum(0).cpu(), ls.grad, rtol=2e-4, atol=2e-4)
    assert s[_lib.STAT_ROWS] == rows
    np.testing.assert_allclose(s[_lib.STAT_LOSS], per_row.sum().item(), rtol=1e-5)
    np.testing.assert_allclose(s[_lib.STAT_KL], (old[idx] - new).sum().item(), rtol=1e-5)
    np.testing.assert_allclose(s[_lib.STAT_ENTROPY], entropy.sum().item(), rtol=1e-5)
    np.testing.assert_allclose(s[_lib.STAT_STD], dist.stddev.sum().item(), rtol=1e-5)
    assert s[_lib.STAT_CLIPPED] == clipped.item()
    assert s[_lib.STAT_NONZERO_ADV] == (adv[idx] != 0).sum().item()


def test_mse_loss(K):
    from tonic_b200 import _lib
    rows, pool = 130, 300
    v, t = torch.randn(rows), torch.randn(pool)
    idx = torch.randint(0, pool, (rows,))
    dout = torch.zeros(rows, 4, device='cuda')
    stats = torch.zeros(_lib.STAT_COUNT, dtype=torch.float64, device='cuda')
    K.mse_loss(v.cuda(), t.cuda(), idx.cuda(), rows, dout, stats)
    np.testing.assert_allclose(dout[:, 0].cpu(), 2 * (v - t[idx]), rtol=1e-6, atol=1e-6)
    s = stats.cpu().numpy()
    np.testing.assert_allclose(s[_lib.STAT_LOSS], ((v - t[idx]) ** 2).sum().item(), rtol=1e-5)
    np.testing.assert_allclose(s[_lib.STAT_VALUE], v.sum().item(), rtol=1e-5, atol=1e-5)
    assert s[_lib.STAT_ROWS] == rows


def test_device_permutation_is_a_bijection(K):
    for n in (1, 2, 7, 1000, 524288, 100003):
        out = torch.empty(n, dtype=torch.int64, device='cuda')
        K.permutation(12345, 0, out)
        host = out.cpu().numpy()
        np.testing.assert_array_equal(np.sort(host), np.arange(n))
        if n >= 1000:
            other = torch.empty_like(out)
            K.permutation(12345, 1, other)
            assert (other.cpu().numpy() != host).mean() > 0.99
            # no structure left: position and value are uncorrelated
            assert abs(np.corrcoef(np.arange(n), host)[0, 1]) < 0.05
            # minibatch slices cover the index range evenly
            chunk = host[:n // 8]
            assert abs(chunk.mean() / n - 0.5) < 0.05


@pytest.mark.parametrize('name,steps', [('Pendulum-v1', 450), ('MountainCarContinuous-v0', 1200)])
@pytest.mark.parametrize('time_feature', [False, True])
def test_classic_control_device_env_is_bit_exact_vs_numpy(name, steps, time_feature):
    """Gym(name) as a device kernel (csrc/classic_env.cu) against the numpy
    restatement (environments/classic.py) under the reference's Sequential semantics
    (distributed.py:28-58): observations, rewards, resets, terminations bit-exact, including
    time-outs, terminations and the counter-based reset stream."""
    import tonic_b200
    from tonic_b200.environments import classic, host
    N, seed = 37, 11
    build = lambda: classic.Gym(name, time_feature=time_feature)     # noqa: E731
    dev = tonic_b200.environments.distribute(build, 1, N)
    assert isinstance(dev, tonic_b200.environments.DeviceVectorEnvironment)
    ref = host.distribute_host(build, 1, N)
    dev.initialize(seed=seed)
    ref.initialize(seed=seed)
    assert dev.max_episode_steps == ref.max_episode_steps
    np.testing.assert_array_equal(dev.start(host=True), ref.start())
    rs = np.random.RandomState(3)
    resets = terms = 0
    for t in range(steps):
        # a slowly varying push (MountainCar needs a consistent strategy to reach the goal)
        a = (np.sin(0.05 * t + np.arange(N))[:, None] * 1.3 + 0.2 * rs.normal(size=(N, 1))).astype(np.float32)
        o1, i1 = dev.step(a)
        o2, i2 = ref.step(a)
        np.testing.assert_array_equal(o1, o2, err_msg=f'step {t}')
        for k in ('observations', 'rewards', 'resets', 'terminations'):
            np.testing.assert_array_equal(i1[k], i2[k], err_msg=f'{k} step {t}')
        resets += int(i1['resets'].sum())
        terms += int(i1['terminations'].sum())
    assert resets >= N                        # every environment hit its time limit / goal at least once
    if name.startswith('MountainCar'):
        assert terms > 0                      # true terminations were exercised
