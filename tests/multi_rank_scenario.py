"""Run under torchrun (one rank per GPU): the workers of a golden scenario are
sharded over the ranks; every rank must reproduce the SINGLE-PROCESS reference's
losses (gradient / statistics all-reduce, global advantage normalisation, global
observation statistics) and end with identical weights.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_rank_scenario.py ppo_small a2c_small
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import scenarios  # noqa: E402
from tests import product, test_gpu_agents  # noqa: E402


def check_all():
    """Every golden scenario whose workers split over the ranks of the initialised process
    group (bench.py --gpus N > 1 runs this before timing: `multi_rank_parity`)."""
    world = dist.get_world_size()
    names = [n for n in ('ppo_small', 'ppo_wide', 'ppo_ragged', 'a2c_small', 'td3_small', 'sac_small',
                         'ddpg_small') if scenarios.SCENARIOS[n]['workers'] % world == 0]
    assert names, f'no golden scenario splits over {world} ranks'
    check(names, verbose=False)
    return names


def main(names):
    local_rank = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    check(names)
    dist.destroy_process_group()


def check(names, verbose=True):
    rank, world = dist.get_rank(), dist.get_world_size()
    from tonic_b200.utils import logger
    saved = {k: getattr(logger, k) for k in ('store', 'store_aggregate')}
    for name in names:
        cfg = scenarios.SCENARIOS[name]
        g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')))
        rec = scenarios.InfoRecorder()
        agent, env = product.build(cfg, log=rec)
        n_local = cfg['workers'] // world
        assert env.workers == n_local and env.first_worker == rank * n_local
        rows = slice(rank * n_local, (rank + 1) * n_local)
        actions = product.teacher_forced(agent, env, g, cfg['vector_steps'], rows=rows)
        if rank == 0 and os.environ.get('TB_DEBUG'):
            got = test_gpu_agents.by_key(rec.keys, rec.means)
            ref = test_gpu_agents.by_key(g['info_keys'], g['info_mean'])
            for k in sorted(ref):
                print(k, 'ours', np.array(got.get(k, []))[:6], 'ref', np.array(ref[k])[:6], flush=True)
        test_gpu_agents.check_infos(rec, g)
        np.testing.assert_allclose(actions, g['actions'][:, rows], rtol=1e-5, atol=5e-5)
        test_gpu_agents.check_weights(agent, g)
        # replicas stay in lock-step
        flat = torch.cat([n.params for n in agent.model.networks()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), 'replica weights diverged'
        if rank == 0 and verbose:
            print(f'{name}: {world} ranks reproduce the single-process reference', flush=True)
    for k, v in saved.items():
        setattr(logger, k, v)


if __name__ == '__main__':
    main(sys.argv[1:] or ['ppo_small'])
