"""CUDA-graph capture of kernel sequences.

The update of one PPO iteration is ~6,000 kernel launches with static shapes; the
rollout of a segment ~600.  Issued one by one through ctypes the host cannot keep
the GPU busy, so (config.graphs) each section is captured once into a CUDA graph
-- after two eager executions that size every workspace -- and then replayed with a
single launch.  Everything that changes between replays lives in device memory:
the segment / parameters / optimizer state, the Philox and permutation stream
positions (tb_counter_add), the KL early-stop flag, the statistics blocks.
"""

import torch

from . import _lib, config

# kernel launches executed through graph replays (the library's own counter,
# tb_launch_count, only sees launches issued -- or captured -- through its entry points)
replayed_launches = 0


class CapturedSection:
    def __init__(self, fn, warmup=2):
        self.fn, self.warmup = fn, warmup
        self.calls, self.graph = 0, None
        self.launches = 0          # kernel nodes captured from this library

    def __call__(self):
        global replayed_launches
        if not config.graphs:
            return self.fn()
        if self.graph is not None:
            self.graph.replay()
            replayed_launches += self.launches
            return
        self.calls += 1
        if self.calls <= self.warmup:
            return self.fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(graph):
            self.fn()
        self.launches = _lib.launch_count() - before      # counted at capture, executed per replay
        self.graph = graph
        graph.replay()        # capture does not execute: run the section once
        replayed_launches += self.launches
