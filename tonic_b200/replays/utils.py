"""Replay helpers with the reference's names (tonic/replays/utils.py)."""

import torch

from .. import kernels


def lambda_returns(values, next_values, rewards, resets, terminations, discount_factor,
                   trace_decay):
    """Reverse scan of utils.py:4-19 on the device ([T, N] float32 arrays)."""
    args = [kernels.to_device(a) for a in (values, next_values, rewards, resets, terminations)]
    out = torch.empty_like(args[2])
    kernels.lambda_returns(*args, out, discount_factor, trace_decay)
    return out


def flatten_batch(values):
    """[T, N, ...] -> [T*N, ...] (utils.py:22-25): a view, flat index t*N+n."""
    return values.reshape((values.shape[0] * values.shape[1],) + tuple(values.shape[2:]))
