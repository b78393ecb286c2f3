import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
torch.set_num_threads(1)
from oracle import bench_shape, port, scenarios
g = dict(np.load('/root/repo/tests/golden/ppo_bench.npz'))
rec = scenarios.InfoRecorder()
agent, env = port.build(bench_shape.CFG, log=rec)
out = bench_shape.drive(agent, env, bench_shape.CFG)
def by_key(keys, vals):
    d = {}
    for k, v in zip(keys, vals): d.setdefault(str(k), []).append(float(v))
    return d
got, ref = by_key(rec.keys, rec.means), by_key(g['info_keys'], g['info_mean'])
for k in ref:
    a, b = np.array(got[k]), np.array(ref[k])
    if a.shape != b.shape: print(k, 'shape', a.shape, b.shape); continue
    err = np.abs(a - b)
    print(k, 'first-epoch abs %.3g rel %.3g | whole abs %.3g rel %.3g' % (err[:32].max(), (err[:32]/(np.abs(b[:32])+1e-12)).max(), err.max(), (err/(np.abs(b)+1e-12)).max()))
w = bench_shape.weight_digests(agent.state_dict(), 'digest_w/')
print('weights max abs diff of first-8 samples', max(np.abs(v[2:] - g[k][2:]).max() for k, v in w.items()))
