"""Multi-GPU: one process per GPU, ciphertext-batch sharding (BASELINE.json configs[2]).

Homomorphic ops on different ciphertexts are independent, so the batch is dealt to ranks in
contiguous slices and every rank runs the same fused calls on its slice: there is NO collective on
the data path.  Keys are replicated once (each rank generates or loads the same key; `broadcast_key`
ships rank 0's copy with torch.distributed).  The reference's alternatives were TensorPipe RPC
(tiberate/extension/multigpu.py) and per-op limb sharding inside one engine with peer copies
(tiberate/ckks_engine.py:1248-1265) -- neither is reproduced.
"""

from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of `total` ciphertexts owned by `rank`; sizes differ by <= 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, rank: int, world: int):
    """View of the rank's slice of a [B, ...] tensor (no copy)."""
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def broadcast_key(parts, src: int = 0, group=None):
    """Replicate a key-switch key (list of (b, a) tensors or None) from `src` to every rank, in place."""
    import torch.distributed as dist

    for part in parts:
        if part is None:
            continue
        for t in part:
            dist.broadcast(t, src=src, group=group)
    return parts


def gather_batch(local, total: int, group=None):
    """All ranks' slices concatenated in rank order (verification only: not on the op path)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
