"""The N > 1 path on the CPU: world_size-2 `gloo` ranks shard a frame batch the way bench.py / a multi-GPU caller does
(contiguous blocks, no data-path collective — SURVEY.md 8e) and the concatenated per-frame results equal the
single-process ones.  The per-frame work here is the product's host stage (it needs no device)."""
import hashlib
import os
import socket

import numpy as np
import pytest


def test_shard_range_partitions():
    from aruco3_b200.sharding import shard_range
    for n in (0, 1, 7, 256, 1024, 1023):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _frame_result(mask):
    import aruco3_b200 as a3
    q = a3.quads_from_mask(mask)
    return len(q), hashlib.sha256(q.tobytes()).hexdigest()


def _worker(rank, world, port, masks_path, out_path):
    import torch.distributed as dist
    from aruco3_b200.sharding import merge_counts, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    masks = np.load(masks_path)
    lo, hi = shard_range(len(masks), rank, world)
    mine = [_frame_result(masks[i]) for i in range(lo, hi)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)  # results only; the frames themselves never cross ranks
    if rank == 0:
        np.save(out_path, np.array(merge_counts(gathered), dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(oracle, tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    from aruco3_b200 import synth
    frames, _ = synth.render_batch("C1", 5)
    masks = np.stack([oracle.adaptive_threshold(oracle.to_luma8(f), 7) for f in frames])
    want = [_frame_result(m) for m in masks]
    masks_path, out_path = tmp_path / "masks.npy", tmp_path / "out.npy"
    np.save(masks_path, masks)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(masks_path), str(out_path)), nprocs=2, join=True)
    got = [tuple(r) for r in np.load(out_path, allow_pickle=True)]
    assert got == want and sum(n for n, _ in got) >= 20
