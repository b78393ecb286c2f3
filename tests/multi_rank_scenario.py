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
    check_fused_exchange(verbose=False)
    return names


def _fast_mode_weights(fused, iterations=3):
    """A few PPO iterations at the benchmarked network shape (17 -> 256 -> 256 -> 6, tanh) in the
    fast mode (device noise, device permutations: every rank owns batch_size / world rows of every
    minibatch), with the gradient exchange inside the fused weight-gradient kernel (`fused`) or
    through the publish / pull kernel pair.  Returns the final flat weights of this rank."""
    import tonic_b200
    import tonic_b200.torch
    from tonic_b200 import config, kernels
    m, n = tonic_b200.torch.models, tonic_b200.torch.normalizers
    world = dist.get_world_size()
    saved = config.noise, config.indices, kernels._PEER_FUSED
    config.noise = config.indices = 'device'
    kernels._PEER_FUSED = fused
    try:
        spec = tonic_b200.environments.SynthControl('HalfCheetah', max_episode_steps=100)
        env = tonic_b200.environments.distribute(lambda: spec, 1, 256 * world)
        env.initialize(seed=3)
        model = m.ActorCritic(
            actor=m.Actor(encoder=m.ObservationEncoder(), torso=m.MLP((256, 256), torch.nn.Tanh),
                          head=m.DetachedScaleGaussianPolicyHead()),
            critic=m.Critic(encoder=m.ObservationEncoder(), torso=m.MLP((256, 256), torch.nn.Tanh),
                            head=m.ValueHead()),
            observation_normalizer=n.MeanStd())
        replay = tonic_b200.replays.Segment(size=32, batch_iterations=3, batch_size=2048 * world)
        agent = tonic_b200.torch.agents.PPO(model=model, replay=replay)
        agent.initialize(env.observation_space, env.action_space, seed=3)
        env.start()
        for _ in range(iterations):
            assert agent.rollout(env, 32) == 32
        torch.cuda.synchronize()
        return torch.cat([net.params for net in agent.model.networks()]).clone()
    finally:
        config.noise, config.indices, kernels._PEER_FUSED = saved


def check_fused_exchange(verbose=True):
    """the gradient exchange fused into the weight-gradient kernel (push over NVLink
    peer memory, csrc/peers.cuh) gives the same bits as the publish / pull kernels (both sum the
    ranks' flat gradients in rank order), and the replicas stay in lock-step."""
    from tonic_b200.utils import logger
    saved = {k: getattr(logger, k) for k in ('store', 'store_aggregate')}
    logger.store = logger.store_aggregate = lambda *a, **k: None
    try:
        fused = _fast_mode_weights(True)
        pulled = _fast_mode_weights(False)
        # the reduce-scatter + all-gather form ranks > 2 use (read per launch from the environment)
        os.environ['TONIC_B200_PEER_TWO_PHASE'] = '1'
        try:
            two_phase = _fast_mode_weights(True)
        finally:
            del os.environ['TONIC_B200_PEER_TWO_PHASE']
    finally:
        for k, v in saved.items():
            setattr(logger, k, v)
    assert torch.isfinite(fused).all()
    ref = fused.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(fused, ref), 'replica weights diverged (fused exchange)'
    diff = (fused - pulled).abs().max().item()
    assert torch.equal(fused, pulled), f'fused exchange differs from the publish / pull kernels: {diff}'
    assert torch.equal(two_phase, pulled), 'two-phase exchange differs from the publish / pull kernels'
    if dist.get_rank() == 0 and verbose:
        print(f'fused exchange: {dist.get_world_size()} ranks, bit-identical to the publish / pull '
              'kernels, replicas in lock-step', flush=True)


def main(names):
    import faulthandler
    # a rank stuck in a collective would hold the GPUs until the caller's timeout: dump where and exit
    faulthandler.dump_traceback_later(int(os.environ.get('TB_MULTI_WATCHDOG', '420')), exit=True)
    local_rank = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    check(names)
    check_fused_exchange()
    # every check passed.  Tearing NCCL down after CUDA graphs captured its collectives (fast mode)
    # can block inside destroy_process_group: give it a few seconds, then leave anyway.
    import gc
    import threading
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    bail = threading.Timer(15.0, lambda: os._exit(0))
    bail.daemon = True
    bail.start()
    dist.destroy_process_group()
    bail.cancel()


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
