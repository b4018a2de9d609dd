"""The N>1 path on CPU: world_size-2 (and 3) gloo process groups exercising the B-scan partition and the final host
gather with a stand-in for the per-rank CUDA call (a deterministic per-B-scan digest), plus partition properties."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT  # noqa: F401

from fdoct_b200 import shard


def test_partition_properties():
    for nB in (0, 1, 7, 8, 125, 1000):
        for world in (1, 2, 3, 4, 8):
            parts = shard.partition(nB, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == nB
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert shard.partition(1000, 8) == [(125 * r, 125 * (r + 1)) for r in range(8)]  # config C4
    assert shard.frame_range((3, 5), 8) == (24, 40)


def _fake_process(frames, A, D, oph):
    """Stand-in for ctx.process_bscans: per B-scan, a digest image that depends on every frame of the unit."""
    nb = frames.shape[0] // A
    out = np.empty((nb, D, oph), np.uint8)
    for b in range(nb):
        s = frames[b * A:(b + 1) * A].astype(np.uint64).sum(axis=0)  # [h, w]
        out[b] = (s[:oph, :D].T % 251).astype(np.uint8)
    return out


def _worker(rank, world, port, nB, A, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 65535, size=(nB * A, 6, 16), dtype=np.uint16)  # same on every rank
    out = shard.process_sharded(lambda f: _fake_process(f, A, 8, 6), frames, A, (8, 6), rank=rank, world=world)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nB,A", [(2, 7, 2), (3, 5, 1), (2, 1, 4)])
def test_sharded_gather_matches_single_process(world, nB, A):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nB, A, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 65535, size=(nB * A, 6, 16), dtype=np.uint16)
    assert np.array_equal(got, _fake_process(frames, A, 8, 6))


def test_frame_source_shards_like_an_array():
    """FrameSource (frames that are not one array in this process, e.g. bench.py's per-rank pinned shard) partitions exactly like
    an ndarray: same slices asked for, same result for world = 1."""
    import numpy as np

    from fdoct_b200 import shard

    frames = np.arange(12 * 2 * 3, dtype=np.uint16).reshape(12, 2, 3)
    asked = []

    def getter(lo, hi):
        asked.append((lo, hi))
        return frames[lo:hi]

    def fn(fr):
        return np.full((fr.shape[0] // 2, 4, 2), fr[0, 0, 0] % 251, np.uint8)

    a = shard.process_sharded(fn, frames, 2, (4, 2), rank=0, world=1)
    b = shard.process_sharded(fn, shard.FrameSource(12, getter), 2, (4, 2), rank=0, world=1)
    assert np.array_equal(a, b) and asked == [(0, 12)]
    for world in (2, 3, 5):
        for r, (lo, hi) in enumerate(shard.partition(6, world)):
            asked.clear()
            shard.FrameSource(12, getter)[slice(*shard.frame_range((lo, hi), 2))]
            assert asked == [(2 * lo, 2 * hi)]
