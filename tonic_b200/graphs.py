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

from . import config


class CapturedSection:
    def __init__(self, fn, warmup=2):
        self.fn, self.warmup = fn, warmup
        self.calls, self.graph = 0, None

    def __call__(self):
        if not config.graphs:
            return self.fn()
        if self.graph is not None:
            self.graph.replay()
            return
        self.calls += 1
        if self.calls <= self.warmup:
            return self.fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.fn()
        self.graph = graph
        graph.replay()        # capture does not execute: run the section once
