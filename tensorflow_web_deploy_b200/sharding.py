"""Data-parallel sharding of a request batch over the GPUs of one box.

Every image is independent at inference (frozen BN statistics, ops.py:122-125), so the path shards
by rows: rank r owns rows [lo, hi) of the batch, runs its own engine (weights replicated, 48.6 MB),
and ONE all-gather of the [rows, C] float32 probabilities assembles the result on every rank — the
only collective on the path (NCCL over NVLink on the B200 box, gloo in the CPU tests).  Batches that
fit one GPU's sweet spot stay on one GPU (north_star: shard "only when the incoming batch exceeds
one GPU's sweet spot").

Two entry points:

* ``classify(x_host)``: host batch in, host probabilities out (the request path; CPU-testable with gloo and
  any ``forward_fn``).
* ``step_device(x_rows_dev)``: this rank's rows already resident on its GPU (the throughput path,
  ``bench.py --gpus N``): the tower is enqueued on the engine's stream, the all-gather of its [rows, C]
  probabilities runs on a side stream and so overlaps the NEXT step's tower (double-buffered outputs);
  nothing touches the host.  ``.wait()`` on the returned handle orders the caller's stream after the gather.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

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


class GatherHandle:
    """Result of ShardedTower.step_device: `probs` [world * rows, C] on the device, valid once `event` has fired."""

    def __init__(self, probs, event):
        self.probs, self.event = probs, event

    def wait(self, stream=None):
        import torch
        (stream or torch.cuda.current_stream()).wait_event(self.event)
        return self.probs

    def synchronize(self):
        self.event.synchronize()
        return self.probs


class ShardedTower:
    """`forward_fn(x_rows) -> probs[rows, C]` (numpy) serves classify(); `engine` (engine.Engine) serves
    step_device().  `dist` is torch.distributed (already initialised) or None for one process."""

    def __init__(self, forward_fn: Optional[Callable[[np.ndarray], np.ndarray]], num_classes: int, dist=None, device=None,
                 sweet_spot: int = SWEET_SPOT, engine=None, depth: int = 2):
        self.forward_fn, self.C, self.dist, self.device, self.sweet_spot = forward_fn, num_classes, dist, device, sweet_spot
        self.engine = engine
        # step_device keeps `depth` output / gather buffers in rotation: a rank may run that many steps ahead of the slowest
        # one before it has to wait for a gather.  A GatherHandle stays valid for depth - 1 further steps.  MEASURED on 2 GPUs
        # (30-step runs): depth 2 2.301 / 2.264 ms per step, 4 2.254, 8 2.350 / 2.277, one GPU 2.267 / 2.238 — all inside the
        # run-to-run spread of power-capped boards, so the default stays 2.
        self.depth = max(2, int(depth))
        if forward_fn is None and engine is not None:
            self.forward_fn = engine.forward
        self._bufs = None
        self._step = 0

    @property
    def world(self) -> int:
        return self.dist.get_world_size() if self.dist is not None else 1

    @property
    def rank(self) -> int:
        return self.dist.get_rank() if self.dist is not None else 0

    # ------------------------------------------------------------------ request path (host in / host out)
    def classify(self, x: np.ndarray) -> np.ndarray:
        import torch
        n = x.shape[0]
        world, rank = self.world, self.rank
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

    # ------------------------------------------------------------------ throughput path (device resident)
    def _device_state(self, rows: int):
        import torch
        if self._bufs is None or self._bufs["rows"] != rows:
            dev = self.device
            self._bufs = {
                "rows": rows,
                "probs": [torch.empty((rows, self.C), dtype=torch.float32, device=dev) for _ in range(self.depth)],
                "gathered": [torch.empty((rows * self.world, self.C), dtype=torch.float32, device=dev) for _ in range(self.depth)],
                "tower_done": [torch.cuda.Event() for _ in range(self.depth)],
                "gather_done": [torch.cuda.Event() for _ in range(self.depth)],
                "comm": torch.cuda.Stream(device=dev),
                "busy": [False] * self.depth,
            }
        return self._bufs

    def step_device(self, x_rows_dev, stream=None) -> GatherHandle:
        """x_rows_dev: float32 (normalised) or uint8 (0..255) [rows, S, S, 3] CUDA tensor holding THIS rank's rows (every
        rank passes the same number of rows).  Enqueues the tower on `stream` (default: torch's current stream) and the all-gather on the
        side stream; returns at once."""
        import torch
        if self.engine is None:
            raise RuntimeError("step_device needs an engine")
        rows = int(x_rows_dev.shape[0])
        st = self._device_state(rows)
        i = self._step % self.depth
        self._step += 1
        stream = stream or torch.cuda.current_stream()
        if st["busy"][i]:
            stream.wait_event(st["gather_done"][i])             # the gather that last read probs[i] has finished
        probs = st["probs"][i]
        self.engine.forward_device(x_rows_dev.data_ptr(), rows, probs.data_ptr(), stream.cuda_stream,
                                   u8=(x_rows_dev.dtype == torch.uint8))
        if self.world == 1:
            st["tower_done"][i].record(stream)
            st["busy"][i] = False
            return GatherHandle(probs, st["tower_done"][i])
        st["tower_done"][i].record(stream)
        comm = st["comm"]
        comm.wait_event(st["tower_done"][i])
        with torch.cuda.stream(comm):
            self.dist.all_gather_into_tensor(st["gathered"][i], probs)      # [rows, C] x world: the only collective
            st["gather_done"][i].record(comm)
        st["busy"][i] = True
        return GatherHandle(st["gathered"][i], st["gather_done"][i])
