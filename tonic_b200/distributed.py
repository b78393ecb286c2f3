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


class PeerRegion:
    """Symmetric (NVLink peer-mapped) buffer of one network for the fused gradient
    all-reduce + Adam kernels (csrc/optim.cu: tb_peer_publish / tb_adam_step_peers).
    Creating it is a collective: every rank must construct its regions in the same
    order (they are created on the first update of each network)."""

    def __init__(self, n_params, fused=False):
        """fused: the push-model layout the fused weight-gradient kernel exchanges through
        (csrc/peers.cuh) instead of the publish / pull pair's."""
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib, kernels
        dev = kernels.device()
        lib = _lib.load()
        nbytes = int(lib.tb_peer_region_bytes_fused(n_params) if fused
                     else lib.tb_peer_region_bytes(n_params))
        self.buffer = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=dev)
        self.buffer.zero_()
        if hasattr(symm_mem, 'enable_symm_mem_for_group'):
            try:
                symm_mem.enable_symm_mem_for_group(dist.group.WORLD.group_name)
            except Exception:
                pass
        self.handle = symm_mem.rendezvous(self.buffer, dist.group.WORLD)
        pointers = list(self.handle.buffer_ptrs)
        assert len(pointers) == world() <= 8, pointers
        self.struct = _lib.TbPeers(world=world(), rank=rank(),
                                   base=(ctypes.c_void_p * 8)(*pointers))
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self.block_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier()          # every region is zeroed before anybody raises a flag
