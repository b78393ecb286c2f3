"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: worker sharding,
partition of the global minibatch / replay index streams into per-rank local
indices, and the sum all-reduce wrapper."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tonic_b200 import distributed, environments
        from tonic_b200.utils.random_state import RandomState
        assert distributed.world() == world and distributed.rank() == rank

        # 1. env sharding: contiguous blocks, np.split order (distributed.py:137)
        spec = environments.SynthControl('HalfCheetah')
        env = environments.distribute(lambda: spec, 2, 6)       # 12 workers in total
        assert env.workers == 6 and env.first_worker == 6 * rank

        # 2. all-reduce wrapper
        t = torch.full((5,), float(rank + 1), dtype=torch.float64)
        distributed.all_reduce(t)
        assert torch.equal(t, torch.full((5,), 3.0, dtype=torch.float64))

        # 3. global permutation -> local rows (Segment layout: index = t * N + n)
        T, n_local = 8, 6
        n_global = n_local * world
        rs = RandomState(5)
        order = np.arange(T * n_global)
        rs.shuffle(order)
        local, mine = distributed.local_rows(order, n_global, n_local, rank)
        batch = 32
        cuts = list(range(0, len(order), batch))
        counts = np.add.reduceat(mine.astype(np.int64), cuts)
        np.save(os.path.join(out_dir, f'local_{rank}.npy'), local)
        np.save(os.path.join(out_dir, f'counts_{rank}.npy'), counts)
        np.save(os.path.join(out_dir, f'order_{rank}.npy'), order)
        # every rank must have drawn the same global stream
        gathered = [torch.zeros(len(order), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(order))
        assert all(torch.equal(g, gathered[0]) for g in gathered)
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    T, n_local = 8, 6
    n_global = n_local * world
    order = np.load(tmp_path / 'order_0.npy')
    locals_ = [np.load(tmp_path / f'local_{r}.npy') for r in range(world)]
    counts = [np.load(tmp_path / f'counts_{r}.npy') for r in range(world)]
    assert sum(len(x) for x in locals_) == len(order)
    # rebuild every global minibatch from the per-rank local rows
    offsets = [0] * world
    for j, lo in enumerate(range(0, len(order), 32)):
        want = sorted(order[lo:lo + 32].tolist())
        got = []
        for r in range(world):
            part = locals_[r][offsets[r]:offsets[r] + counts[r][j]]
            offsets[r] += counts[r][j]
            t, n = part // n_local, part % n_local
            got += (t * n_global + r * n_local + n).tolist()
        assert sorted(got) == want
        assert sum(c[j] for c in counts) == len(want)


def test_local_rows_single_rank_identity():
    from tonic_b200 import distributed
    idx = np.random.RandomState(0).permutation(40)
    local, mine = distributed.local_rows(idx, 5, 5, 0)
    np.testing.assert_array_equal(local, idx)
    assert mine.all()
