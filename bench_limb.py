#!/usr/bin/env python
"""bench_limb.py -- BASELINE.json configs[3]: logN=17 deep-modulus key switching with the RNS limbs sharded
over the N GPUs of one box (one process per GPU; per key switch ONE NCCL all-gather of the ModUp digits over
NVLink, issued asynchronously and hidden under the ModUp of the digit groups the rank owns and under the
kernels of the previous key switch of the stream).

Job: a stream of B independent logN17 polynomials (73 ordinary + 6 special limbs at level 0, 13 digit groups)
is key switched per step -- STRONG scaling: the job is fixed, the limbs are dealt to the ranks.  In the same
run every rank checks its rows against the unsharded key switch of the same inputs (bit-exact), and rank 0
times the unsharded batched key switch as the 1-GPU figure the speed-up is quoted against.

Used by bench.py (`--mode limb`, and as the `limb_sharded` block of the default line when N > 1), or
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_limb.py
"""

from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def run(rank: int, world: int, local: int, steps: int = 10, warmup: int = 3, batch: int = 8, level: int = 0,
        logN: int = 17, overlap: bool = True, slices: int = 1, shard_special: bool = True) -> dict | None:
    import torch
    import torch.distributed as dist

    from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context
    from tiberate_fhe_b200.dist import LimbShardedKeySwitch
    from tiberate_fhe_b200.presets import PRESETS

    dev = torch.device("cuda", local)
    q, K = PRESETS[logN]["q"], PRESETS[logN]["K"]
    N, P, no = 1 << logN, len(q), len(q) - K
    gen = torch.Generator(device=dev).manual_seed(0xB217)  # the SAME inputs on every rank

    def uniform(primes, *lead):
        t = torch.empty(*lead, len(primes), N, dtype=torch.int64, device=dev)
        for i, qi in enumerate(primes):
            t[..., i, :].random_(0, int(qi), generator=gen)
        return t

    full = Tb200Context(logN, q, K, device=local)
    full.set_chunk(batch)
    ng = full.num_groups0
    key_full = [(uniform(q), uniform(q)) for _ in range(ng)]
    a = uniform(q[level:no], batch)  # [B, L, N]
    L = no - level
    r0, r1 = torch.empty_like(a), torch.empty_like(a)
    kv_full = KeySwitchKeyView(key_full, N)
    full.keyswitch(level, a, kv_full, r0, r1)  # reference result of this run

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps

    ms_one = None
    if rank == 0:  # the 1-GPU figure, measured while the other ranks wait at the barrier inside timed()
        o0, o1 = torch.empty_like(a), torch.empty_like(a)
    ms_full = timed((lambda: full.keyswitch(level, a, kv_full, o0, o1)) if rank == 0 else (lambda: None))
    if world == 1:
        full.close()
        return {"metric": f"key-switch ops/s at logN={logN}", "value": batch / (ms_full / 1e3), "unit": "ops/s",
                "n_gpus": 1, "ms_per_step": ms_full, "batch": batch, "scaling": "strong",
                "config": {"workload": f"logN{logN} preset, level {level}: {L} ordinary + {K} special limbs, {ng} digit "
                                       f"groups; {batch} polynomials key switched per step (unsharded, one batched call)"}}
    ctx = Tb200Context(logN, q, K, device=local, rank=rank, world=world)
    ctx.set_chunk(batch)
    ids = ctx.local_prime_ids
    rows = [g - level for g in ctx.local_rows(level)]
    key = KeySwitchKeyView([(b[ids].contiguous(), a_[ids].contiguous()) for b, a_ in key_full], N)
    a_loc = a[:, rows].contiguous()
    o0l, o1l = torch.empty_like(a_loc), torch.empty_like(a_loc)
    ks = LimbShardedKeySwitch(ctx, overlap=overlap, shard_special=shard_special)

    nsl = max(1, min(slices, batch))
    cuts = [(i * batch // nsl, (i + 1) * batch // nsl) for i in range(nsl)]

    # The batch is a stream of `slices` batched key switches.  Software pipeline of depth one, across step boundaries:
    # the digits + all-gather of the NEXT slice are issued (tb200_ks_digits + ncclAllGather, on their two alternating
    # state slots) before the current slice is finished, so the collective travels under the current slice's kernels.
    pend = {"st": None, "k": 0}

    def step(last=True):  # last: nothing follows (drain): every started key switch is finished inside the call
        for i, (lo, hi) in enumerate(cuts):
            if pend["st"] is None:
                pend["st"] = ks.start(level, a_loc[lo:hi], slot=pend["k"] % 2)
            cur = pend["st"]
            pend["st"] = None
            nxt = None
            if not (last and i == nsl - 1):
                nlo, nhi = cuts[(i + 1) % nsl]

                def nxt(nlo=nlo, nhi=nhi):
                    pend["st"] = ks.start(level, a_loc[nlo:nhi], slot=(pend["k"] + 1) % 2)

            ks.finish(level, cur, key, o0l[lo:hi], o1l[lo:hi], between=nxt)
            pend["k"] += 1

    def timed_stream():
        for i in range(warmup):
            step(last=i == warmup - 1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(last=i == steps - 1)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps

    step()
    torch.cuda.synchronize()
    ok = bool(torch.equal(o0l, r0[:, rows]) and torch.equal(o1l, r1[:, rows]))
    o0l.zero_()
    o1l.zero_()
    step(last=False)  # and through the pipelined path: two steps, the second one drains
    step(last=True)
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(o0l, r0[:, rows]) and torch.equal(o1l, r1[:, rows]))
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ms = timed_stream()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    t0 = torch.tensor([ms_full if rank == 0 else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t0, op=dist.ReduceOp.MAX)
    ms_one = float(t0.item())
    S = ctx.ks_state_info(level)[0]
    out = None
    if rank == 0:
        # limb count only (a special limb costs ~2.5 scale limbs, so the replicated flow sits below its figure)
        sp_mode = "their key sums sharded too (second all-gather of 2 K N words per polynomial)" if ks.shard_special \
            else "special limbs replicated"
        ideal = (L + K) / (-(-L // world) + (-(-K // world) if ks.shard_special else K))
        out = {"metric": f"key-switch ops/s at logN={logN}, limb-sharded", "value": batch / (ms / 1e3), "unit": "ops/s",
               "n_gpus": world, "ms_per_step": ms, "batch": batch, "scaling": "strong",
               "one_gpu_ops_per_s": batch / (ms_one / 1e3), "speedup_vs_one_gpu": ms_one / ms,
               "ideal_speedup": ideal, "bit_exact_vs_unsharded": bool(int(flag.item())), "overlap": overlap, "slices": nsl,
               "shard_special": bool(ks.shard_special),
               "allgather_bytes_per_keyswitch": S * N * 8 + (2 * ctx.ks_sp_info()[0] * N * 8 if ks.shard_special else 0),
               "config": {"workload": f"logN{logN} preset, level {level}: {L} ordinary + {K} special limbs, {ng} digit "
                                      f"groups; {batch} polynomials key switched per step, limbs dealt to {world} ranks by "
                                      f"digit group, {sp_mode} (limb-count ideal {ideal:.2f})",
                          "local_limbs_rank0": len(rows)}}
    if not bool(int(flag.item())):
        raise SystemExit("limb-sharded key switch differs from the unsharded one")
    ctx.close()
    full.close()
    return out


def main():
    import argparse

    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--slices", type=int, default=1)
    ap.add_argument("--replicate-special", action="store_true", help="the reference's flow: special limbs on every rank")
    args, _ = ap.parse_known_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run(rank, world, local, args.steps, args.warmup, args.batch, args.level, overlap=not args.no_overlap, slices=args.slices,
              shard_special=not args.replicate_special)
    if rank == 0:
        out.update({"steps": args.steps, "warmup": args.warmup, "dtype": "int64", "data": "synthetic",
                    "higher_is_better": True})
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
