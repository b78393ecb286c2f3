"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over
NVLink / NVSwitch; gloo in the CPU tests).

The reference's "distributed" training is P worker processes stepping
environments for ONE learner (tonic/environments/distributed.py:69-155).  Here
environments shard contiguously over the ranks (rank r owns workers
[r*N/W, (r+1)*N/W), the `np.split` order of distributed.py:137), every rank holds
a replica of the weights / optimizer state, and each minibatch update all-reduces
the flat gradient buffer (sum of per-sample gradients) and the statistics block;
dividing by the GLOBAL minibatch size afterwards reproduces the single-process
mean exactly.
"""

import numpy as np
import torch
import torch.distributed as dist


def initialized():
    return dist.is_available() and dist.is_initialized()


def world():
    return dist.get_world_size() if initialized() else 1


def rank():
    return dist.get_rank() if initialized() else 0


def all_reduce(tensor):
    """In-place sum over ranks (no-op for a single process)."""
    if world() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def local_rows(global_indices, workers_global, workers_local, rank_):
    """Splits flat transition indices of the GLOBAL layout (index = row *
    workers_global + worker) into the entries owned by `rank_`, re-expressed in
    its LOCAL layout (row * workers_local + local worker).  Order is preserved."""
    g = np.asarray(global_indices, np.int64)
    row, worker = g // workers_global, g % workers_global
    mine = (worker // workers_local) == rank_
    return row[mine] * workers_local + (worker[mine] - rank_ * workers_local), mine
