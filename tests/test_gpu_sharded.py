"""Limb-sharded key switch over NCCL (BASELINE configs[3]); needs >= 2 GPUs (skipped otherwise):
  gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu
"""

import os
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port):
    import numpy as np
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    import datetime

    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                            timeout=datetime.timedelta(seconds=90))
    try:
        from oracle.context import OracleContext, toy_primes
        from oracle.engine import OracleEngine
        from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context
        from tiberate_fhe_b200.dist import LimbShardedKeySwitch, shard_rows
        from tiberate_fhe_b200.presets import PRESETS

        # 1. toy ring against the oracle
        logN, ns, K = 12, 7, 2
        q = toy_primes(logN, ns, K)
        octx = OracleContext(logN, q, K)
        eng = OracleEngine(octx)
        rng = np.random.default_rng(17)
        sk, _ = eng.gen_secret(rng)
        evk = eng.gen_evk(rng, sk)
        ctx = Tb200Context(logN, q, K, device=rank, rank=rank, world=world)
        ks = LimbShardedKeySwitch(ctx)
        ids = ctx.local_prime_ids
        key = KeySwitchKeyView([(torch.from_numpy(np.ascontiguousarray(p[0][ids])).to(dev),
                                 torch.from_numpy(np.ascontiguousarray(p[1][ids])).to(dev)) for p in evk], octx.N)
        for level in (0, 3, 6):
            a = eng.uniform(rng, octx.level_primes(level, False))
            want0, want1 = eng.create_switcher(a, evk, level)
            a_loc = shard_rows(torch.from_numpy(a), ctx, level).to(dev)
            o0, o1 = torch.zeros_like(a_loc), torch.zeros_like(a_loc)
            ks(level, a_loc, key, o0, o1)
            rows = [g - level for g in ctx.local_rows(level)]
            assert np.array_equal(o0.cpu().numpy(), want0[rows]), (rank, level)
            assert np.array_equal(o1.cpu().numpy(), want1[rows]), (rank, level)
        # limb-sharded rescale: one NCCL broadcast of the dropped limb, then local kernels
        from tiberate_fhe_b200.dist import LimbShardedRescale

        rs = LimbShardedRescale(ctx)
        for level in (0, 2, 5):
            lp = octx.level_primes(level, False)
            ct = [eng.uniform(rng, lp), eng.uniform(rng, lp)]
            want = eng.rescale(ct, level)
            loc = [shard_rows(torch.from_numpy(x), ctx, level).to(dev) for x in ct]
            r0, r1 = rs(level, loc[0], loc[1])
            rows = [g - level - 1 for g in ctx.local_rows(level + 1)]
            assert np.array_equal(r0.cpu().numpy(), want[0][rows]), (rank, level)
            assert np.array_equal(r1.cpu().numpy(), want[1][rows]), (rank, level)
        # whole operations on sharded ciphertexts: rotate_single and cc_mult + relinearize against the oracle
        from tiberate_fhe_b200.context import galois_element
        from tiberate_fhe_b200.dist import LimbShardedOps

        delta = 3
        rotk = eng.gen_rotk(rng, sk, delta)
        rotk_loc = KeySwitchKeyView([(torch.from_numpy(np.ascontiguousarray(p[0][ids])).to(dev),
                                      torch.from_numpy(np.ascontiguousarray(p[1][ids])).to(dev)) for p in rotk], octx.N)
        for overlap in (True, False):
            ops = LimbShardedOps(ctx, overlap=overlap)
            for level in (0, 2, 5):
                lp = octx.level_primes(level, False)
                ct1 = [eng.uniform(rng, lp), eng.uniform(rng, lp)]
                ct2 = [eng.uniform(rng, lp), eng.uniform(rng, lp)]
                l1 = [shard_rows(torch.from_numpy(x), ctx, level).to(dev) for x in ct1]
                l2 = [shard_rows(torch.from_numpy(x), ctx, level).to(dev) for x in ct2]
                want = eng.rotate_single(ct1, rotk, delta, level)
                o0, o1 = ops.rotate(level, galois_element(octx.N, delta), l1[0], l1[1], rotk_loc)
                rows = [g - level for g in ctx.local_rows(level)]
                assert np.array_equal(o0.cpu().numpy(), want[0][rows]), ("rotate", rank, level, overlap)
                assert np.array_equal(o1.cpu().numpy(), want[1][rows]), ("rotate", rank, level, overlap)
                want, lvl1 = eng.cc_mult(ct1, ct2, evk, level, pre_rescale=True)
                o0, o1 = ops.cc_mult_relin(level, l1[0], l1[1], l2[0], l2[1], key)
                rows = [g - lvl1 for g in ctx.local_rows(lvl1)]
                assert np.array_equal(o0.cpu().numpy(), want[0][rows]), ("cc_mult_relin", rank, level, overlap)
                assert np.array_equal(o1.cpu().numpy(), want[1][rows]), ("cc_mult_relin", rank, level, overlap)
        ctx.close()

        # 2. logN17 (79 primes, 13 digit groups): sharded result == unsharded result of rank 0
        q, K = PRESETS[17]["q"], PRESETS[17]["K"]
        P, N, no = len(q), 1 << 17, len(q) - K
        level = 40
        gen = torch.Generator(device=dev).manual_seed(99)  # same stream on every rank

        def uniform(primes):
            t = torch.empty(len(primes), N, dtype=torch.int64, device=dev)
            for i, qi in enumerate(primes):
                t[i].random_(0, int(qi), generator=gen)
            return t

        ng = -(-(no - 1) // K) + 1
        full_key = [(uniform(q), uniform(q)) for _ in range(ng)]
        a = uniform(q[level:no])
        ctx = Tb200Context(17, q, K, device=rank, rank=rank, world=world)
        ks = LimbShardedKeySwitch(ctx)
        ids = ctx.local_prime_ids
        key = KeySwitchKeyView([(b[ids].contiguous(), a_[ids].contiguous()) for b, a_ in full_key], N)
        a_loc = shard_rows(a, ctx, level)
        o0, o1 = torch.zeros_like(a_loc), torch.zeros_like(a_loc)
        ks(level, a_loc, key, o0, o1)
        rows = [g - level for g in ctx.local_rows(level)]
        ctx.close()
        ref = Tb200Context(17, q, K, device=rank)
        r0, r1 = torch.zeros_like(a), torch.zeros_like(a)
        ref.keyswitch(level, a, KeySwitchKeyView(full_key, N), r0, r1)
        assert torch.equal(o0, r0[rows]) and torch.equal(o1, r1[rows]), rank
        ref.close()
    finally:
        dist.destroy_process_group()


def test_limb_sharded_keyswitch_nccl():
    import torch
    import torch.multiprocessing as mp

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)
