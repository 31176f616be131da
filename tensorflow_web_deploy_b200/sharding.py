"""Data-parallel sharding of a request batch over the GPUs of one box.

Every image is independent at inference (frozen BN statistics, ops.py:122-125), so the path shards
by rows: rank r owns rows [lo, hi) of the batch, runs its own engine (weights replicated, 48.6 MB),
and ONE all-gather of the [rows, C] float32 probabilities assembles the result on every rank — the
only collective on the path (NCCL over NVLink on the B200 box, gloo in the CPU tests).  Batches that
fit one GPU's sweet spot stay on one GPU (north_star: shard "only when the incoming batch exceeds
one GPU's sweet spot").
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np

SWEET_SPOT = 128     # images per B200 at which the tower saturates (BASELINE configs[2]/[3])


def gpus_for_batch(n: int, world: int, sweet_spot: int = SWEET_SPOT) -> int:
    """How many ranks take part: ceil(n / sweet_spot), clamped to [1, world]."""
    if n <= 0:
        return 1
    return max(1, min(world, -(-n // sweet_spot)))


def shard_bounds(n: int, parts: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced rows of `rank` among `parts` (first n % parts ranks get one extra row);
    ranks >= parts get the empty range."""
    if rank >= parts:
        return n, n
    base, extra = divmod(n, parts)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedTower:
    """forward_fn(x_rows) -> probs[rows, C] (numpy) for this rank's rows; classify() returns the full
    [n, C] on every rank.  `dist` is torch.distributed (already initialised) or None for one process."""

    def __init__(self, forward_fn: Callable[[np.ndarray], np.ndarray], num_classes: int, dist=None, device=None,
                 sweet_spot: int = SWEET_SPOT):
        self.forward_fn, self.C, self.dist, self.device, self.sweet_spot = forward_fn, num_classes, dist, device, sweet_spot

    def classify(self, x: np.ndarray) -> np.ndarray:
        import torch
        n = x.shape[0]
        world = self.dist.get_world_size() if self.dist is not None else 1
        rank = self.dist.get_rank() if self.dist is not None else 0
        parts = gpus_for_batch(n, world, self.sweet_spot)
        lo, hi = shard_bounds(n, parts, rank)
        mine = self.forward_fn(x[lo:hi]) if hi > lo else np.zeros((0, self.C), np.float32)
        if world == 1:
            return np.asarray(mine, np.float32)
        rows = -(-n // parts)                                   # padded shard height, equal on all ranks
        buf = torch.zeros((rows, self.C), dtype=torch.float32, device=self.device)
        if hi > lo:
            buf[:hi - lo] = torch.as_tensor(np.asarray(mine, np.float32)).to(buf.device)
        out = torch.empty((world * rows, self.C), dtype=torch.float32, device=self.device)
        self.dist.all_gather_into_tensor(out, buf)              # the path's single collective
        out = out.cpu().numpy().reshape(world, rows, self.C)
        pieces = []
        for r in range(parts):
            a, b = shard_bounds(n, parts, r)
            pieces.append(out[r, :b - a])
        return np.concatenate(pieces, 0)
