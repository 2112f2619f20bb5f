"""Frame-batch sharding across GPUs: frames are independent — `detect` is a pure function of one image
(/root/reference/src/aruco.rs:52-121) — so a batch is cut into contiguous blocks, one per rank, and no collective
is needed on the data path; results are concatenated in frame order."""
from __future__ import annotations


def shard_range(n_frames: int, rank: int, world_size: int) -> tuple:
    """Frames [lo, hi) of rank `rank`: contiguous blocks, sizes differ by at most one, earlier ranks get the extras."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, extra = divmod(n_frames, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_counts(per_rank: list) -> list:
    """Concatenate per-rank per-frame result lists in rank order (= global frame order)."""
    out = []
    for part in per_rank:
        out.extend(part)
    return out


class ShardedDetector:
    """One `Detector` per listed device (a device may be listed more than once), each on its own host thread over its
    contiguous block of frames; the per-frame results come back in frame order.  ctypes releases the GIL during
    `a3_detect_batch`, so the shards really run side by side.  The in-process counterpart of `bench.py --gpus N`'s one
    process per GPU, and the Python twin of `aruco3::ShardedDetector` (include/aruco3_b200.hpp)."""

    def __init__(self, config=None, dictionary="ARUCO", devices=(0,), **kw):
        from .detector import Detector
        if not devices:
            raise ValueError("ShardedDetector: no devices")
        self.shards = [Detector(config, dictionary, device=dev, **kw) for dev in devices]

    def detect_batch(self, frames, **kw) -> list:
        import threading
        world = len(self.shards)
        parts, errors = [None] * world, [None] * world

        def run(rank):
            lo, hi = shard_range(len(frames), rank, world)
            try:
                parts[rank] = self.shards[rank].detect_batch(frames[lo:hi], **kw) if hi > lo else []
            except BaseException as e:  # re-raised in the caller's thread
                errors[rank] = e

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e
        return merge_counts(parts)

    def close(self):
        for s in self.shards:
            s.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
