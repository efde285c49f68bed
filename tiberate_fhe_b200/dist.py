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
def prime_owner(ctx, p: int) -> int:
    """Group rank that owns ordinary prime `p` (global id) under the limb partition of
    tiberate/context/rns_partition.py:34-52 as libtb200 applies it (csrc/tb200.cu: tb_group_owner):
    scale-prime group g (K consecutive primes) -> rank (np - 1 - g) mod world, the base prime -> rank 0."""
    K = ctx.K
    ns = ctx.P_global - K - 1
    np_ = -(-ns // K)
    return (np_ - 1 - p // K) % ctx.world if p < ns else 0


class LimbShardedKeySwitch:
    """Key switch of a polynomial whose limbs are dealt to ranks by digit group
    (tiberate/context/rns_partition.py:34-52; the reference moves the digit states with per-device
    `tensor.to(device)` copies, ckks_engine.py:1248-1265).

    Per call: each rank computes the ModUp digits of the groups it owns (tb200_ks_digits), ONE all-gather
    completes the digit-state buffer on every rank (its layout is owner-major with equal segments, so the
    gather is in place and copy-free), then each rank extends / transforms / multiplies / ModDowns its own
    limbs.  The special limbs are replicated, as in the reference, so ModDown needs no second exchange.

    overlap=True (default): the all-gather is issued asynchronously and the ModUp of the groups this rank
    owns (tb200_ks_modup, which=1: extend + forward pass A, ~1/world of a third of the key switch) runs
    while it is in flight; the other groups' ModUp (which=2) and the core (tb200_ks_core) follow once the
    digits have arrived.  overlap=False: gather, then tb200_ks_finish."""

    def __init__(self, ctx, group=None, overlap: bool = True, shard_special: bool = True):
        import torch.distributed as dist

        self.ctx, self.group, self.overlap = ctx, group, overlap
        self.world = dist.get_world_size(group)
        if ctx.world != self.world or ctx.rank != dist.get_rank(group):
            raise ValueError("context rank/world must match the process group")
        self._state = {}
        # shard_special (needs overlap=True and the mod-q path): the key sums of the K special limbs are sharded too --
        # every rank extends / transforms / multiplies only its share of them (they are 60-bit limbs, 2.5x the cost
        # of a scale limb, and replicated they cap the speed-up at ~1.7x / 2.5x on 2 / 4 GPUs), and a second, small
        # all-gather (2 K N words per polynomial) completes them before ModDown while the ordinary limbs are still
        # being processed.  Same bits as the replicated flow (include/tb200.h: tb200_ks_core_sp).
        self.shard_special = bool(shard_special and overlap and self.world > 1)
        self._sp = {}

    def _levels_all_ranks_busy(self, level: int) -> bool:
        """True when every rank still owns an ordinary limb at `level` (the sharded-special flow assumes it; deep
        levels where a rank has run out of limbs use the replicated flow -- decided by rule, the same on every rank)."""
        ctx = self.ctx
        owners = {prime_owner(ctx, p) for p in range(level, ctx.P_global - ctx.K)}
        return len(owners) == self.world

    def _sp_buffer(self, batch, like, slot):
        import torch

        rows, seg, _, _ = self.ctx.ks_sp_info()
        key = (batch, like.device, slot)
        sp = self._sp.get(key)
        if sp is None:
            sp = torch.zeros((rows, batch, 2, self.ctx.N), dtype=torch.int64, device=like.device)
            self._sp[key] = sp
        return sp, seg

    def _state_buffer(self, level, like, slot=0):
        """Digit-state buffer, stored [state_rows, (batch,) N]: the rows of one owner are contiguous for the whole
        batch, so a batched key switch still needs ONE in-place all-gather.  Returns (storage, view handed to the
        library, segment row0, segment rows); the view is [batch, state_rows, N] with strides (N, batch N, 1)."""
        import torch

        S, row0, seg, _ = self.ctx.ks_state_info(level)
        B = like.shape[0] if like.dim() == 3 else 0
        key = (level, like.device, slot, B)
        st = self._state.get(key)
        if st is None:
            st = torch.zeros((S, B, self.ctx.N) if B else (S, self.ctx.N), dtype=torch.int64, device=like.device)
            self._state[key] = st
        return st, (st.permute(1, 0, 2) if B else st), row0, seg

    def start(self, level: int, a_local, slot: int = 0):
        """Digits of the owned groups + the all-gather (asynchronous with overlap=True).  Several key switches
        can be started back to back on different slots: their collectives then run under the ModUp / core
        kernels of the earlier ones (every rank must start and finish them in the same order)."""
        import torch.distributed as dist

        store, state, row0, seg = self._state_buffer(level, a_local, slot)
        if a_local.shape[-2] > 0:  # deep levels can leave a rank without ordinary limbs
            self.ctx.ks_digits(level, a_local, state)
        work = None
        if self.world > 1:  # every rank takes part, also those that own nothing at this level
            views = [store[r * seg:(r + 1) * seg] for r in range(self.world)]
            work = dist.all_gather(views, store[row0:row0 + seg], group=self.group, async_op=self.overlap)
        return state, (work if self.overlap else None), slot

    def finish(self, level: int, started, ksk_local, out0, out1, add0=None, add1=None, tail: int = 0, between=None):
        """between: optional callable that starts the NEXT key switch of a stream (`start(...)` on another slot).
        It is invoked where its collective cannot delay this one: collectives of one communicator run in issue
        order, so the next key switch's (large) digit all-gather must be issued after this one's special-limb
        all-gather, not before."""
        ctx = self.ctx
        state, work, slot = started
        sharded_sp = self.shard_special and self._levels_all_ranks_busy(level)
        if between is not None and not sharded_sp:
            between()
        if out0.shape[-2] == 0:
            if work is not None:
                work.wait()
            return out0, out1
        if not self.overlap:
            ctx.ks_finish(level, state, ksk_local, out0, out1, add0=add0, add1=add1, tail=tail)
            return out0, out1
        if sharded_sp:
            import torch.distributed as dist

            batch = out0.shape[0] if out0.dim() == 3 else 1
            sp, seg = self._sp_buffer(batch, out0, slot)
            ctx.ks_modup(level, state, which=1 + 4)
            if work is not None:
                work.wait()
            ctx.ks_modup(level, state, which=2 + 4)
            ctx.ks_core_sp(level, batch, ksk_local, sp)  # this rank's share of the special limbs first ...
            views = [sp[r * seg:(r + 1) * seg] for r in range(self.world)]
            work2 = dist.all_gather(views, sp[ctx.rank * seg:(ctx.rank + 1) * seg], group=self.group, async_op=True)
            if between is not None:
                between()
            ctx.ks_core_ord(level, batch, ksk_local, out0)  # ... the ordinary limbs while those sums travel
            work2.wait()
            ctx.ks_moddown(level, sp, out0, out1, add0=add0, add1=add1, tail=tail)
            return out0, out1
        ctx.ks_modup(level, state, which=1)  # own groups: their digits never left this GPU
        if work is not None:
            work.wait()  # the compute stream waits for the collective; the host does not block
        ctx.ks_modup(level, state, which=2)
        ctx.ks_core(level, state, ksk_local, out0, out1, add0=add0, add1=add1, tail=tail)
        return out0, out1

    def __call__(self, level: int, a_local, ksk_local, out0, out1, add0=None, add1=None, tail: int = 0):
        """a_local / out*: this rank's rows [L_local, N] or [batch, L_local, N] (coefficient domain, canonical;
        batch <= the context's chunk); ksk_local: KeySwitchKeyView over the LOCAL key rows ([P_local, N] per group)."""
        return self.finish(level, self.start(level, a_local), ksk_local, out0, out1, add0, add1, tail)


class LimbShardedRescale:
    """Rescale of a limb-sharded ciphertext (tiberate/ckks_engine.py:1520-1618; the reference ships the
    dropped limb to the other devices with `tensor.to(device)`, :1560-1575).

    The rank that owns prime `level` broadcasts the dropped limb of both polynomials (2 N int64 -- the
    one exchange of this operation; the owner follows from the partition rule, no lookup collective), then
    every rank rescales the limbs it keeps with the op-layer kernel (tb200_rescale_rows).  Returns this
    rank's rows at level + 1 (fresh tensors).  The per-level scale tables are built once and cached."""

    def __init__(self, ctx, group=None):
        import torch.distributed as dist

        self.ctx, self.group = ctx, group
        self.world = dist.get_world_size(group)
        if ctx.world != self.world or ctx.rank != dist.get_rank(group):
            raise ValueError("context rank/world must match the process group")
        self._scales = {}

    def _scale_table(self, level, kept, device):
        import torch

        key = (level, str(device))
        t = self._scales.get(key)
        if t is None:
            ctx, R = self.ctx, 1 << 62
            ql = ctx.q_global[level]
            t = torch.tensor([(pow(ql, -1, ctx.q_global[g]) * R) % ctx.q_global[g] for g in kept], dtype=torch.int64,
                             device=device)
            self._scales[key] = t
        return t

    def __call__(self, level: int, c0_local, c1_local, exact: bool = True):
        import torch
        import torch.distributed as dist

        ctx = self.ctx
        ids = ctx.local_rows(level)
        owner_rank = prime_owner(ctx, level)
        owner = owner_rank == ctx.rank
        assert owner == (level in ids)
        drop = torch.empty(2, ctx.N, dtype=torch.int64, device=c0_local.device)
        if owner:  # the dropped prime is the smallest alive id: local row 0
            drop[0].copy_(c0_local[0])
            drop[1].copy_(c1_local[0])
        if self.world > 1:
            src = dist.get_global_rank(self.group, owner_rank) if self.group is not None else owner_rank
            dist.broadcast(drop, src=src, group=self.group)
        kept = [g for g in ids if g > level]
        outs = []
        for i, c in enumerate((c0_local, c1_local)):
            out = (c[1:] if owner else c).clone()
            if kept:
                ctx.rescale_rows(out, ctx.local_prime_ids.index(kept[0]), self._scale_table(level, kept, c.device),
                                 drop[i], ctx.q_global[level] // 2, exact)
            outs.append(out)
        return outs[0], outs[1]


class LimbShardedOps:
    """Whole homomorphic operations on limb-sharded ciphertexts (BASELINE.json configs[3]): every step other
    than the two exchanges above is limb-local and runs on this rank's rows only.

      rotate(level, galois, c0, c1, rotk)      tiberate/ckks_engine.py:1804-1840: automorphism (local) +
                                                key switch of the rotated c1 with the switch-key tail
      cc_mult_relin(level, a, b, evk)           :1640-1732: rescale both operands (one broadcast each), tensor
                                                product in the NTT domain (local), inverse transforms (local),
                                                key switch of d2 with the relinearisation tail
    Results are this rank's rows of the reference's result, bit for bit (tests/test_dist_gloo.py,
    tests/test_gpu_sharded.py)."""

    def __init__(self, ctx, group=None, overlap: bool = True):
        self.ctx = ctx
        self.ks = LimbShardedKeySwitch(ctx, group, overlap)
        self.rs = LimbShardedRescale(ctx, group)

    def rotate(self, level: int, galois: int, c0, c1, rotk_local):
        import torch

        r0, r1 = torch.empty_like(c0), torch.empty_like(c1)
        self.ctx.rotate(level, galois, c0, c1, None, r0, r1)  # automorphism only: limb-local
        o0, o1 = torch.empty_like(c0), torch.empty_like(c1)
        self.ks(level, r1, rotk_local, o0, o1, add0=r0, tail=2)
        return o0, o1

    def cc_mult_relin(self, level: int, a0, a1, b0, b1, evk_local):
        import torch

        ctx = self.ctx
        x0, x1 = self.rs(level, a0, a1)
        y0, y1 = (x0, x1) if (a0 is b0 and a1 is b1) else self.rs(level, b0, b1)
        lvl = level + 1
        d = [torch.empty_like(x0) for _ in range(3)]
        if x0.shape[-2] > 0:
            ctx.cc_mult_triplet(lvl, x0, x1, y0, y1, d[0], d[1], d[2], False)
            p0 = ctx.local_prime_ids.index(ctx.local_rows(lvl)[0])
            for t in d:  # intt_radix2_exit_reduce (ckks_engine.py:1709-1711)
                ctx.intt(t, p0, 2)
        o0, o1 = torch.empty_like(x0), torch.empty_like(x1)
        self.ks(lvl, d[2], evk_local, o0, o1, add0=d[0], add1=d[1], tail=1)
        return o0, o1


def shard_rows(t, ctx, level: int, with_special: bool = False):
    """Rows of a full [L(+K), N] level-`level` tensor that live on ctx's rank (a copy, contiguous)."""
    rows = [g - level for g in ctx.local_rows(level)]
    if with_special:
        n_ord = ctx.P_global - ctx.K - level
        rows += list(range(n_ord, n_ord + ctx.K))
    return t[rows].contiguous()
