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


# ------------------------------------------------------------------------------------------------
# RNS-limb sharding for the largest rings (BASELINE.json configs[3])
# ------------------------------------------------------------------------------------------------
class LimbShardedKeySwitch:
    """Key switch of a polynomial whose limbs are dealt to ranks by digit group
    (tiberate/context/rns_partition.py:34-52; the reference moves the digit states with per-device
    `tensor.to(device)` copies, ckks_engine.py:1248-1265).

    Per call: each rank computes the ModUp digits of the groups it owns (tb200_ks_digits), ONE
    all-gather completes the digit-state buffer on every rank (its layout is owner-major with equal
    segments, so the gather is in place and copy-free), then each rank extends / transforms /
    multiplies / ModDowns its own limbs (tb200_ks_finish).  The special limbs are replicated, as in
    the reference, so ModDown needs no second exchange.
    """

    def __init__(self, ctx, group=None):
        import torch.distributed as dist

        self.ctx, self.group = ctx, group
        self.world = dist.get_world_size(group)
        if ctx.world != self.world or ctx.rank != dist.get_rank(group):
            raise ValueError("context rank/world must match the process group")
        self._state = {}

    def _state_buffer(self, level, like):
        import torch

        S, row0, seg, _ = self.ctx.ks_state_info(level)
        key = (level, like.device)
        st = self._state.get(key)
        if st is None:
            st = torch.zeros(S, self.ctx.N, dtype=torch.int64, device=like.device)
            self._state[key] = st
        return st, row0, seg

    def __call__(self, level: int, a_local, ksk_local, out0, out1, add0=None, add1=None, tail: int = 0):
        """a_local / out*: this rank's rows [L_local, N] (coefficient domain, canonical);
        ksk_local: KeySwitchKeyView over the LOCAL key rows ([P_local, N] per digit group)."""
        import torch.distributed as dist

        state, row0, seg = self._state_buffer(level, out0)
        has_rows = out0.shape[-2] > 0  # deep levels can leave a rank without ordinary limbs
        if has_rows:
            self.ctx.ks_digits(level, a_local, state)
        if self.world > 1:  # every rank takes part, also those that own nothing at this level
            views = [state[r * seg:(r + 1) * seg] for r in range(self.world)]
            dist.all_gather(views, state[row0:row0 + seg], group=self.group)
        if has_rows:
            self.ctx.ks_finish(level, state, ksk_local, out0, out1, add0=add0, add1=add1, tail=tail)
        return out0, out1


class LimbShardedRescale:
    """Rescale of a limb-sharded ciphertext (tiberate/ckks_engine.py:1520-1618; the reference ships the
    dropped limb to the other devices with `tensor.to(device)`, :1560-1575).

    The rank that owns prime `level` broadcasts the dropped limb of both polynomials (2 N int64 -- the
    one exchange of this operation), then every rank rescales the limbs it keeps with the op-layer kernel
    (tb200_rescale_rows).  Returns this rank's rows at level + 1 (fresh tensors)."""

    def __init__(self, ctx, group=None):
        import torch.distributed as dist

        self.ctx, self.group = ctx, group
        self.world = dist.get_world_size(group)
        if ctx.world != self.world or ctx.rank != dist.get_rank(group):
            raise ValueError("context rank/world must match the process group")

    def __call__(self, level: int, c0_local, c1_local, exact: bool = True):
        import torch
        import torch.distributed as dist

        ctx = self.ctx
        ids = ctx.local_rows(level)
        owner = level in ids
        drop = torch.zeros(2, ctx.N, dtype=torch.int64, device=c0_local.device)
        if owner:  # the dropped prime is the smallest alive id: local row 0
            drop[0].copy_(c0_local[0])
            drop[1].copy_(c1_local[0])
        if self.world > 1:
            who = torch.tensor([ctx.rank if owner else -1], dtype=torch.int64, device=c0_local.device)
            dist.all_reduce(who, op=dist.ReduceOp.MAX, group=self.group)
            dist.broadcast(drop, src=int(who.item()), group=self.group)
        kept = [g for g in ids if g > level]
        outs = []
        for i, c in enumerate((c0_local, c1_local)):
            out = (c[1:] if owner else c).clone()
            if kept:
                ql, R = ctx.q_global[level], 1 << 62
                scales = torch.tensor([(pow(ql, -1, ctx.q_global[g]) * R) % ctx.q_global[g] for g in kept],
                                      dtype=torch.int64, device=c.device)
                ctx.rescale_rows(out, ctx.local_prime_ids.index(kept[0]), scales, drop[i], ql // 2, exact)
            outs.append(out)
        return outs[0], outs[1]


def shard_rows(t, ctx, level: int, with_special: bool = False):
    """Rows of a full [L(+K), N] level-`level` tensor that live on ctx's rank (a copy, contiguous)."""
    rows = [g - level for g in ctx.local_rows(level)]
    if with_special:
        n_ord = ctx.P_global - ctx.K - level
        rows += list(range(n_ord, n_ord + ctx.K))
    return t[rows].contiguous()
