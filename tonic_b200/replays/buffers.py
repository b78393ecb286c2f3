"""Off-policy ring replay resident in HBM (reference: tonic/replays/buffers.py).

Storage is the reference's layout -- per key a float32 array [max_size, N, ...]
with max_size = size // N (buffers.py:40-45), written at row `index`, flat sample
index = row * N + worker -- so `RandomState(seed).randint(size * N)` addresses
the same transitions (buffers.py:84-89).  Batches are not materialised: the update
kernels gather rows through the index vector.  `discounts = (1 - terminations) *
discount_factor` (buffers.py:34-36) is recomputed inside the target kernel from
the stored terminations (bit-identical float32 arithmetic).
With n-step returns (`return_steps > 1`, buffers.py:58-79) the discounts column is
stored too and `tb_replay_accumulate_n_steps` back-fills the previous rows at every store.
"""

import numpy as np
import torch

from .. import distributed, kernels
from .._lib import ptr
from ..utils.random_state import RandomState


class Buffer:
    def __init__(self, size=int(1e6), return_steps=1, batch_iterations=50, batch_size=100,
                 discount_factor=0.99, steps_before_batches=int(1e4), steps_between_batches=50):
        self.full_max_size = size
        self.return_steps = return_steps
        self.batch_iterations = batch_iterations
        self.batch_size = batch_size
        self.discount_factor = discount_factor
        self.steps_before_batches = steps_before_batches
        self.steps_between_batches = steps_between_batches

    def initialize(self, seed=None):
        self.np_random = RandomState(seed)       # buffers.py:22
        self.buffers = None
        self.index = 0
        self.size = 0
        self.last_steps = 0

    def ready(self, steps):                      # buffers.py:28-31
        if steps < self.steps_before_batches:
            return False
        return (steps - self.last_steps) >= self.steps_between_batches

    def allocate(self, **shapes):
        dev = kernels.device()
        self.num_workers = next(iter(shapes.values()))[0]
        self.max_size = self.full_max_size // self.num_workers
        self.buffers = {k: torch.full((self.max_size,) + tuple(s), float('nan'),
                                      dtype=torch.float32, device=dev)
                        for k, s in shapes.items()}
        E, B = self.batch_iterations, self.batch_size
        self._host_idx = torch.empty(E, B, dtype=torch.int64).pin_memory()
        self._dev_idx = torch.empty(E, B, dtype=torch.int64, device=dev)
        # device mirror of (index, size, size * workers) for the graph-safe fast path
        self._ring = torch.zeros(3, dtype=torch.int64, device=dev)
        self._ring_tables = {}
        self._index_counter = kernels.new_counter()

    def store(self, **kwargs):
        if self.buffers is None:
            self.allocate(**{k: tuple(v.shape) if isinstance(v, torch.Tensor) else np.shape(v)
                             for k, v in kwargs.items()})
        for key, val in kwargs.items():
            kernels.to_device(val, out=self.buffers[key][self.index])
        if self.return_steps > 1:
            self.accumulate_n_steps()
        self.advance()

    def accumulate_n_steps(self):
        """buffers.py:34-36,58-79: discounts = float32(1 - terminations) * discount_factor for the
        row just written, then the n-step back-fill of the previous rows."""
        b = self.buffers
        if 'discounts' not in b:
            b['discounts'] = torch.full_like(b['rewards'], float('nan'))
        b['discounts'][self.index] = (1 - b['terminations'][self.index]) * np.float32(self.discount_factor)
        kernels.replay_accumulate_n_steps(b['rewards'], b['discounts'], b['next_observations'],
                                          b['resets'], self.index, self.size, self.return_steps)

    # -- device-resident fast path (config.noise == config.indices == 'device') ---------------
    def sync_ring(self):
        """Host -> device copy of the ring position (after host-side stores)."""
        self._ring.copy_(torch.tensor([self.index, self.size, self.size * self.num_workers]))

    def store_device(self, **staged):
        """Buffer.store (buffers.py:47-56) of this vector step's rows, given as device tensors
        with FIXED addresses, at the row the device ring state points to (CUDA-graph safe:
        nothing here depends on the host's copy of `index`)."""
        import ctypes
        keys = tuple(staged)
        table = self._ring_tables.get(keys)
        if table is None or any(staged[k].data_ptr() != p for k, p in zip(keys, table[3])):
            n = len(keys)
            src = (ctypes.c_void_p * n)(*[staged[k].data_ptr() for k in keys])
            dst = (ctypes.c_void_p * n)(*[self.buffers[k].data_ptr() for k in keys])
            elems = (ctypes.c_int64 * n)(*[self.buffers[k][0].numel() for k in keys])
            for k in keys:
                assert staged[k].numel() == self.buffers[k][0].numel() and staged[k].is_contiguous(), k
            table = self._ring_tables[keys] = (src, dst, elems, [staged[k].data_ptr() for k in keys])
        from .. import _lib
        _lib.call('tb_ring_store', table[0], table[1], table[2], len(keys), ptr(self._ring),
                  kernels.stream())

    def advance_device(self):
        from .. import _lib
        _lib.call('tb_ring_advance', ptr(self._ring), self.max_size, self.num_workers, kernels.stream())

    def index_batches_device(self, seed):
        """`batch_iterations` index vectors drawn on the device (Philox) from the filled part of
        THIS rank's ring (each rank contributes batch_size / world rows, like the on-policy
        device mode); yields the tuples of `index_batches`."""
        from .. import _lib
        world = distributed.world()
        rows = self.batch_size // world
        assert rows * world == self.batch_size, 'batch_size must split over the ranks'
        E = self.batch_iterations
        _lib.call('tb_randint', int(seed or 0) ^ 0x1d8, distributed.rank(), ptr(self._index_counter),
                  ptr(self._ring[2:]), E * rows, ptr(self._dev_idx), kernels.stream())
        kernels.counter_add(self._index_counter, E * rows)
        flat = self._dev_idx.view(-1)
        for e in range(E):
            yield flat[e * rows:(e + 1) * rows], rows, self.batch_size, None

    def row(self, key):
        """Row `index` of a buffer, for producers that write it in place."""
        return self.buffers[key][self.index]

    def advance(self, mirror_only=False):
        """`mirror_only`: the device ring state was advanced by `advance_device` (fast path)."""
        self.index = (self.index + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)
        self._ring_stale = not mirror_only

    def flat(self, key):
        b = self.buffers[key]
        return b.view((b.shape[0] * b.shape[1],) + tuple(b.shape[2:]))

    def index_batches(self, steps):
        """Yields `batch_iterations` tuples (idx, rows, rows_global, mine): device int64
        flat LOCAL indices of one batch, drawn like buffers.py:84-88 from the GLOBAL
        ring (`size * N_global`, same seed on every rank); a rank keeps the samples
        whose worker column it owns.  Sets `last_steps` afterwards."""
        world, rank = distributed.world(), distributed.rank()
        total = self.size * self.num_workers * world
        host = self._host_idx.numpy()
        counts, masks = [], []
        torch.cuda.current_stream().synchronize()      # previous use of the pinned block
        for e in range(self.batch_iterations):
            if world == 1:
                self.np_random.randint(total, self.batch_size, out=host[e])
                counts.append(self.batch_size)
                masks.append(None)
            else:
                drawn = self.np_random.randint(total, self.batch_size)
                local, mine = distributed.local_rows(drawn, self.num_workers * world,
                                                     self.num_workers, rank)
                host[e, :len(local)] = local
                counts.append(len(local))
                masks.append(torch.from_numpy(mine))
        self._dev_idx.copy_(self._host_idx, non_blocking=True)
        for e in range(self.batch_iterations):
            # masks[e]: which samples of the global batch this rank owns (used to take
            # the matching rows of host-generated noise); None in a single process
            yield self._dev_idx[e, :counts[e]], counts[e], self.batch_size, masks[e]
        self.last_steps = steps

    def get(self, *keys, steps):
        """Reference-style generator of gathered batches (convenience)."""
        for idx, _, _, _ in self.index_batches(steps):
            out = {}
            for k in keys:
                if k == 'discounts' and self.return_steps == 1:
                    # buffers.py:34-36: (1 - terminations) * discount_factor; with n-step
                    # returns the stored column holds the accumulated discounts (:58-79)
                    out[k] = (1 - self.flat('terminations')[idx]) * np.float32(self.discount_factor)
                else:
                    out[k] = self.flat(k)[idx]
            yield out
