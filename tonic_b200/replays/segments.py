"""On-policy segment replay resident in HBM (reference:
tonic/replays/segments.py:6-78).

Storage is the reference's layout -- one float32 array [T, N, ...] per key, flat
transition index t*N+n -- so minibatch indices drawn from the same
`RandomState(seed)` stream address the same transitions.  lambda-returns and the
advantage normalisation run as kernels (csrc/returns.cu); minibatches are never
materialised: the update kernels gather rows through the index vector.
"""

import numpy as np
import torch

from .. import config, distributed, kernels
from ..utils.random_state import RandomState
from . import utils


class Segment:
    def __init__(self, size=4096, batch_iterations=80, batch_size=None, discount_factor=0.99,
                 trace_decay=0.97):
        self.max_size = size
        self.batch_iterations = batch_iterations
        self.batch_size = batch_size
        self.discount_factor = discount_factor
        self.trace_decay = trace_decay

    def initialize(self, seed=None):
        self.np_random = RandomState(seed)      # segments.py:20
        self.seed = seed
        self.buffers = None
        self.index = 0
        self._pinned = None

    def ready(self):
        return self.index == self.max_size

    # -- storage ------------------------------------------------------------------
    def allocate(self, **shapes):
        """shapes: key -> per-step shape, e.g. observations=(N, O)."""
        dev = kernels.device()
        self.num_workers = next(iter(shapes.values()))[0]
        self.buffers = {k: torch.zeros((self.max_size,) + tuple(s), dtype=torch.float32,
                                       device=dev) for k, s in shapes.items()}
        self._workspace = torch.zeros(4, dtype=torch.float64, device=dev)

    def store(self, **kwargs):
        if self.buffers is None:
            self.allocate(**{k: tuple(np.shape(v)) if not isinstance(v, torch.Tensor)
                             else tuple(v.shape) for k, v in kwargs.items()})
        for key, val in kwargs.items():     # host arrays go straight into the segment row
            kernels.to_device(val, out=self.buffers[key][self.index])
        self.index += 1

    def advance(self):
        """For producers that wrote row `index` of the buffers in place."""
        self.index += 1

    # -- graph-safe store at a DEVICE-resident row index (host-protocol fast path) -----------
    def prepare_device_store(self):
        """Outside any captured section: creates / re-synchronises the device row index when the
        host moved `index` since the last device store (get_full, fused rollouts)."""
        if getattr(self, '_ring', None) is None:
            self._ring = torch.zeros(3, dtype=torch.int64, device=kernels.device())
            self._ring_host = 0
            self._ring_tables = {}
        if self._ring_host != self.index % self.max_size:
            self._ring.copy_(torch.tensor([self.index % self.max_size, self.index, 0]))
            self._ring_host = self.index % self.max_size

    def store_device(self, **staged):
        """Segment.store (segments.py:27-36) of rows staged in device tensors with fixed
        addresses, at row *ring[0], then the device index advances (wraps at T).  CUDA-graph
        safe; `prepare_device_store` before and `note_device_store` after, both on the host."""
        import ctypes
        from .. import _lib
        keys = tuple(staged)
        ptrs = [staged[k].data_ptr() for k in keys]
        table = self._ring_tables.get(keys)
        if table is None or table[3] != ptrs:
            n = len(keys)
            table = self._ring_tables[keys] = (
                (ctypes.c_void_p * n)(*ptrs),
                (ctypes.c_void_p * n)(*[self.buffers[k].data_ptr() for k in keys]),
                (ctypes.c_int64 * n)(*[self.buffers[k][0].numel() for k in keys]), ptrs)
        _lib.call('tb_ring_store', table[0], table[1], table[2], len(keys), _lib.ptr(self._ring),
                  kernels.stream())
        _lib.call('tb_ring_advance', _lib.ptr(self._ring), self.max_size, self.num_workers,
                  kernels.stream())

    def note_device_store(self):
        """Host mirror of one `store_device` (called outside the captured section)."""
        self.index += 1
        self._ring_host = self.index % self.max_size

    # -- returns / advantages -----------------------------------------------------------
    def compute_returns(self, values, next_values):
        shape = self.buffers['rewards'].shape
        self.buffers['values'] = kernels.to_device(values).view(shape)
        self.buffers['next_values'] = kernels.to_device(next_values).view(shape)
        if 'returns' not in self.buffers:
            self.buffers['returns'] = torch.empty(shape, dtype=torch.float32,
                                                  device=self.buffers['rewards'].device)
        b = self.buffers
        kernels.lambda_returns(b['values'], b['next_values'], b['rewards'], b['resets'],
                               b['terminations'], b['returns'], self.discount_factor,
                               self.trace_decay)

    def compute_advantages(self):
        """returns - values, normalised over the WHOLE segment (segments.py:41-46).
        When the workers are sharded over several GPUs the mean / variance are
        all-reduced (two passes, like numpy's std) so every rank normalises with the
        statistics of the global segment."""
        b = self.buffers
        if 'advantages' not in b:
            b['advantages'] = torch.empty_like(b['returns'])
        if distributed.world() == 1:
            kernels.advantages(b['returns'], b['values'], b['advantages'], self._workspace)
            return
        n_global = b['returns'].numel() * distributed.world()
        kernels.advantages(b['returns'], b['values'], b['advantages'], self._workspace,
                           n_global, 1)
        distributed.all_reduce(self._workspace)
        kernels.advantages(b['returns'], b['values'], b['advantages'], self._workspace,
                           n_global, 2)
        distributed.all_reduce(self._workspace[2:3])
        kernels.advantages(b['returns'], b['values'], b['advantages'], self._workspace,
                           n_global, 3)

    def get_full(self, *keys):
        self.index = 0
        if 'advantages' in keys:
            self.compute_advantages()
        return {k: utils.flatten_batch(self.buffers[k]) for k in keys}

    # -- minibatches ------------------------------------------------------------------
    def index_batches(self):
        """Yields (idx, rows, rows_global) per minibatch, following
        segments.py:50-65: per iteration one `RandomState.shuffle` of the running
        permutation of all T*N transitions, cut in contiguous slices of
        `batch_size` (the last may be short).  `idx` is a device int64 vector of
        flat LOCAL transition indices (None = the whole local segment in order);
        with several ranks the permutation is over the GLOBAL segment (same seed on
        every rank) and each rank keeps the entries it owns, so `rows` <=
        `rows_global`."""
        world, rank = distributed.world(), distributed.rank()
        local_total = self.max_size * self.num_workers
        total = local_total * world
        if self.batch_size is None:
            for _ in range(self.batch_iterations):
                yield None, local_total, total
            return
        E = self.batch_iterations
        if config.indices == 'device':
            yield from self._device_index_batches(world, rank, local_total)
            return
        cuts = list(range(0, total, self.batch_size))
        if self._pinned is None or self._pinned.shape != (E, local_total):
            self._pinned = torch.empty(E, local_total, dtype=torch.int64).pin_memory()
            self._device_order = torch.empty(E, local_total, dtype=torch.int64,
                                             device=kernels.device())
        order = np.arange(total)
        host = self._pinned.numpy()
        counts = np.zeros((E, len(cuts)), np.int64)
        torch.cuda.current_stream().synchronize()      # previous use of the pinned block
        for e in range(E):
            self.np_random.shuffle(order)
            if world == 1:
                host[e] = order
                counts[e] = [min(self.batch_size, total - lo) for lo in cuts]
            else:
                local, mine = distributed.local_rows(
                    order, self.num_workers * world, self.num_workers, rank)
                host[e] = local
                counts[e] = np.add.reduceat(mine.astype(np.int64), cuts)
        self._device_order.copy_(self._pinned, non_blocking=True)
        for e in range(E):
            offset = 0
            for j, lo in enumerate(cuts):
                rows = int(counts[e, j])
                yield (self._device_order[e, offset:offset + rows], rows,
                       min(self.batch_size, total - lo))
                offset += rows

    def _device_index_batches(self, world, rank, local_total):
        E = self.batch_iterations
        if self.batch_size % world:
            raise ValueError('batch_size must be divisible by the number of ranks')
        local_batch = self.batch_size // world
        if getattr(self, '_device_perm', None) is None or \
                self._device_perm.shape != (E, local_total):
            self._device_perm = torch.empty(E, local_total, dtype=torch.int64,
                                            device=kernels.device())
            self._perm_counter = kernels.new_counter()      # permutations drawn so far
        seed = ((self.seed or 0) << 8) ^ rank
        for e in range(E):
            kernels.permutation(seed, e, self._device_perm[e], device_counter=self._perm_counter)
        kernels.counter_add(self._perm_counter, E)
        for e in range(E):
            for lo in range(0, local_total, local_batch):
                rows = min(local_batch, local_total - lo)
                yield self._device_perm[e, lo:lo + rows], rows, rows * world

    def get(self, *keys):
        """Reference-style generator of gathered minibatches (convenience; the
        agents use `index_batches` and gather inside the kernels)."""
        batch = self.get_full(*keys)
        for idx, rows, _ in self.index_batches():
            if idx is None:
                yield batch
            else:
                yield {k: v[idx] for k, v in batch.items()}
