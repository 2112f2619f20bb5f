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
