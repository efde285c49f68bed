"""World-size-2 gloo tests of the batch-sharding host logic (tiberate_fhe_b200/dist.py).  The op
path itself has no collective; these check the slicing, key replication and result gathering."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tiberate_fhe_b200.dist import broadcast_key, gather_batch, shard_batch, shard_range


def test_shard_range_partitions_every_batch():
    for total in (0, 1, 7, 8, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a "batch" of ciphertext polynomials [B, limbs, N]; each rank works on its slice only
        full = torch.arange(total * 3 * 8, dtype=torch.int64).reshape(total, 3, 8)
        mine = shard_batch(full, rank, world)
        lo, hi = shard_range(total, rank, world)
        assert mine.shape[0] == hi - lo and torch.equal(mine, full[lo:hi])
        out = mine * 2 + 1  # stand-in for the per-ciphertext op
        gathered = gather_batch(out, total)
        assert torch.equal(gathered, full * 2 + 1)
        # key replication: rank 0's key wins
        key = [(torch.full((4, 8), 7 + rank, dtype=torch.int64), torch.full((4, 8), 9 + rank, dtype=torch.int64)), None]
        broadcast_key(key, src=0)
        assert int(key[0][0][0, 0]) == 7 and int(key[0][1][0, 0]) == 9
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total", [5, 8])
def test_two_rank_batch_sharding_gloo(total):
    mp.spawn(_worker, args=(2, _free_port(), total), nprocs=2, join=True)
