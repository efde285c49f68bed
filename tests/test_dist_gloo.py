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


def _sharded_ks_worker(rank, world, port, seed):
    """Runs the real limb-sharded key-switch orchestration (dist.LimbShardedKeySwitch) with gloo on CPU
    tensors; the kernels are the host-compiled sources of tests/emu (no GPU in this container)."""
    import subprocess

    import numpy as np

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        so = os.path.join(root, "tests", "emu", "_build", "libtb200_emu.so")
        if not os.path.exists(so):
            subprocess.run(["sh", os.path.join(root, "tests", "emu", "build_emu.sh")], check=True, capture_output=True)
        from oracle.context import OracleContext, toy_primes
        from oracle.engine import OracleEngine
        from tiberate_fhe_b200 import _native
        from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context
        from tiberate_fhe_b200.dist import LimbShardedKeySwitch, shard_rows

        logN, ns, K = 8, 5, 2
        q = toy_primes(logN, ns, K)
        octx = OracleContext(logN, q, K)
        eng = OracleEngine(octx)
        rng = np.random.default_rng(seed)  # same stream on every rank
        sk, _ = eng.gen_secret(rng)
        evk = eng.gen_evk(rng, sk)
        ctx = Tb200Context(logN, q, K, lib=_native.Lib(so), rank=rank, world=world)
        ks = LimbShardedKeySwitch(ctx)
        ids = ctx.local_prime_ids
        key = KeySwitchKeyView([None if p is None else (torch.from_numpy(np.ascontiguousarray(p[0][ids])),
                                                        torch.from_numpy(np.ascontiguousarray(p[1][ids]))) for p in evk],
                               octx.N)
        assert ks.shard_special  # default flow: the special limbs' key sums are sharded too (second all-gather)
        ks_repl = LimbShardedKeySwitch(ctx, shard_special=False)  # the reference's replicated special limbs
        for level in (0, 2, 4):  # at level 4 rank 1 owns no ordinary limb any more (both fall back to replication)
            a = eng.uniform(rng, octx.level_primes(level, False))
            want0, want1 = eng.create_switcher(a, evk, level)
            a_loc = shard_rows(torch.from_numpy(a), ctx, level)
            rows = [g - level for g in ctx.local_rows(level)]
            for k_ in (ks, ks_repl):
                o0, o1 = torch.zeros_like(a_loc), torch.zeros_like(a_loc)
                k_(level, a_loc, key, o0, o1)
                assert np.array_equal(o0.numpy(), want0[rows]) and np.array_equal(o1.numpy(), want1[rows]), \
                    (rank, level, k_.shard_special)
        # batched: [B, L_local, N] through one all-gather (state stored [rows, B, N])
        ctx.set_chunk(3)
        for level in (0, 2):
            lp = octx.level_primes(level, False)
            ab = np.stack([eng.uniform(rng, lp) for _ in range(3)])
            rows = [g - level for g in ctx.local_rows(level)]
            a_loc = torch.from_numpy(np.ascontiguousarray(ab[:, rows]))
            o0, o1 = torch.zeros_like(a_loc), torch.zeros_like(a_loc)
            ks(level, a_loc, key, o0, o1)
            for b in range(3):
                w0, w1 = eng.create_switcher(ab[b], evk, level)
                assert np.array_equal(o0[b].numpy(), w0[rows]) and np.array_equal(o1[b].numpy(), w1[rows]), (rank, level, b)
        # limb-sharded rescale: the owner of the dropped prime broadcasts it, every rank rescales its rows
        from tiberate_fhe_b200.dist import LimbShardedRescale

        rs = LimbShardedRescale(ctx)
        for level in (0, 1, 3):
            lp = octx.level_primes(level, False)
            ct = [eng.uniform(rng, lp), eng.uniform(rng, lp)]
            want = eng.rescale(ct, level)
            loc = [shard_rows(torch.from_numpy(x), ctx, level) for x in ct]
            o0, o1 = rs(level, loc[0], loc[1])
            rows = [g - level - 1 for g in ctx.local_rows(level + 1)]
            assert np.array_equal(o0.numpy(), want[0][rows]) and np.array_equal(o1.numpy(), want[1][rows]), (rank, level)
        # whole operations on sharded ciphertexts (dist.LimbShardedOps): rotate_single and cc_mult + relinearize,
        # with and without the overlapped (split-call) key switch
        from tiberate_fhe_b200.context import galois_element
        from tiberate_fhe_b200.dist import LimbShardedOps

        def local_key(k):
            return KeySwitchKeyView([None if p is None else (torch.from_numpy(np.ascontiguousarray(p[0][ids])),
                                                             torch.from_numpy(np.ascontiguousarray(p[1][ids]))) for p in k],
                                    octx.N)

        delta = 3
        rotk = eng.gen_rotk(rng, sk, delta)
        rotk_loc, evk_loc = local_key(rotk), key
        for overlap in (True, False):
            ops = LimbShardedOps(ctx, overlap=overlap)
            for level in (0, 1, 3):
                lp = octx.level_primes(level, False)
                ct1 = [eng.uniform(rng, lp), eng.uniform(rng, lp)]
                ct2 = [eng.uniform(rng, lp), eng.uniform(rng, lp)]
                l1 = [shard_rows(torch.from_numpy(x), ctx, level) for x in ct1]
                l2 = [shard_rows(torch.from_numpy(x), ctx, level) for x in ct2]
                want = eng.rotate_single(ct1, rotk, delta, level)
                o0, o1 = ops.rotate(level, galois_element(octx.N, delta), l1[0], l1[1], rotk_loc)
                rows = [g - level for g in ctx.local_rows(level)]
                assert np.array_equal(o0.numpy(), want[0][rows]) and np.array_equal(o1.numpy(), want[1][rows]), \
                    ("rotate", rank, level, overlap)
                want, lvl1 = eng.cc_mult(ct1, ct2, evk, level, pre_rescale=True)
                o0, o1 = ops.cc_mult_relin(level, l1[0], l1[1], l2[0], l2[1], evk_loc)
                rows = [g - lvl1 for g in ctx.local_rows(lvl1)]
                assert np.array_equal(o0.numpy(), want[0][rows]) and np.array_equal(o1.numpy(), want[1][rows]), \
                    ("cc_mult_relin", rank, level, overlap)
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_limb_sharded_keyswitch_two_ranks_gloo(emu_lib):  # the fixture (re)builds tests/emu before the ranks start
    mp.spawn(_sharded_ks_worker, args=(2, _free_port(), 5), nprocs=2, join=True)
