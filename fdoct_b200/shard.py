"""Sharding of a batch of B-scans across the GPUs of one box: one process per GPU (torchrun), no collective on the
data path, only a final host gather (BASELINE.json north_star).

The unit of work is one OUTPUT B-scan = `averages` consecutive frames: the per-B-scan min-max normalise
(BscanFFT.cpp:1254) is the only cross-A-scan dependency and it stays inside the unit.  Rank r owns the contiguous
block partition(nB, world)[r]; it reconstructs it on its own GPU through the C ABI and the display images are
gathered to rank 0 over a host (gloo) group.  torch.distributed is plumbing only.
"""
from __future__ import annotations

import numpy as np


def partition(n_bscans: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced [start, stop) B-scan ranges; the first n_bscans % world ranks get one extra."""
    if world < 1:
        raise ValueError("world must be >= 1")
    base, extra = divmod(n_bscans, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((s, s + n))
        s += n
    return out


def frame_range(bscan_range: tuple[int, int], averages: int) -> tuple[int, int]:
    return bscan_range[0] * averages, bscan_range[1] * averages


class FrameSource:
    """A batch of frames that is not one array in this process: `getter(lo, hi)` returns frames [lo, hi) (e.g. a slice of a
    memory-mapped volume, or - in bench.py - the rank's own pinned copy of its shard).  Only slicing is supported."""

    def __init__(self, nframes: int, getter):
        self.nframes, self.getter = int(nframes), getter

    def __getitem__(self, sl):
        lo, hi, step = sl.indices(self.nframes)
        if step != 1:
            raise ValueError("contiguous slices only")
        return self.getter(lo, hi)


def process_sharded(process_fn, frames: np.ndarray, averages: int, out_shape_per_bscan: tuple[int, int], *, rank: int,
                    world: int, group=None, dst: int = 0):
    """Run `process_fn(frames_slice) -> uint8 [nb, D, oph]` on this rank's B-scans and gather on `dst`.

    `frames` is the WHOLE batch (every rank sees the same host array / memory map, only its slice is touched);
    `group` must be a host-capable (gloo) process group when world > 1.  Returns the full [nB, D, oph] array on
    `dst`, None elsewhere."""
    nframes = frames.nframes if isinstance(frames, FrameSource) else frames.shape[0]
    nB = nframes // averages
    if nframes != nB * averages:
        raise ValueError("nframes must be a multiple of averages")
    parts = partition(nB, world)
    lo, hi = frame_range(parts[rank], averages)
    mine = process_fn(frames[lo:hi]) if hi > lo else np.empty((0,) + tuple(out_shape_per_bscan), np.uint8)
    if mine.shape != (parts[rank][1] - parts[rank][0],) + tuple(out_shape_per_bscan) or mine.dtype != np.uint8:
        raise ValueError(f"process_fn returned {mine.dtype}{mine.shape}")
    if world == 1:
        return mine
    import torch
    import torch.distributed as dist

    # final host gather: equal-sized padded buffers (ranks may differ by one B-scan)
    cap = max(b - a for a, b in parts)
    buf = torch.zeros((cap,) + tuple(out_shape_per_bscan), dtype=torch.uint8)
    buf[: mine.shape[0]] = torch.from_numpy(np.ascontiguousarray(mine))
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = np.empty((nB,) + tuple(out_shape_per_bscan), np.uint8)
    for r, (a, b) in enumerate(parts):
        out[a:b] = gathered[r][: b - a].numpy()
    return out
