#!/usr/bin/env python
"""bench_sharded.py -- BASELINE.json configs[3]: logN=17 deep-modulus key switching with the RNS limbs
sharded over N GPUs (one process per GPU, one NCCL all-gather of the ModUp digits per key switch).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench_sharded.py [--steps K] [--warmup W] [--level L]
  python bench_sharded.py            # N = 1: the unsharded key switch on one GPU

Strong scaling: ONE logN17 polynomial (73 ordinary + 6 special limbs at level 0, 13 digit groups) is key
switched per step; prints one JSON line on rank 0 (key switches per second, max over ranks, CUDA events).
"""

from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--logN", type=int, default=17)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context
    from tiberate_fhe_b200.dist import LimbShardedKeySwitch
    from tiberate_fhe_b200.presets import PRESETS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    q, K = PRESETS[args.logN]["q"], PRESETS[args.logN]["K"]
    N, no = 1 << args.logN, len(q) - K
    ctx = Tb200Context(args.logN, q, K, device=local, rank=rank, world=world)
    ids = ctx.local_prime_ids
    gen = torch.Generator(device=dev).manual_seed(0xB200 + rank)

    def uniform(primes):
        t = torch.empty(len(primes), N, dtype=torch.int64, device=dev)
        for i, qi in enumerate(primes):
            t[i].random_(0, int(qi), generator=gen)
        return t

    ng = ctx.num_groups0
    key = KeySwitchKeyView([(uniform([q[i] for i in ids]), uniform([q[i] for i in ids])) for _ in range(ng)], N)
    rows = ctx.local_rows(args.level)
    a = uniform([q[g] for g in rows]) if rows else torch.zeros(1, N, dtype=torch.int64, device=dev)
    o0, o1 = torch.zeros_like(a), torch.zeros_like(a)
    if world > 1:
        ks = LimbShardedKeySwitch(ctx)
        step = lambda: ks(args.level, a, key, o0, o1)  # noqa: E731
    else:
        step = lambda: ctx.keyswitch(args.level, a, key, o0, o1)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        L = no - args.level
        S = ctx.ks_state_info(args.level)[0]
        print(json.dumps({
            "metric": f"key-switch ops/s at logN={args.logN}, limb-sharded", "value": args.steps / (ms / 1e3),
            "unit": "ops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "scaling": "strong", "dtype": "int64", "data": "synthetic",
            "config": {"workload": f"logN{args.logN} preset, level {args.level}: {L} ordinary + {K} special limbs, "
                                   f"{ng} digit groups; create_switcher of one polynomial per step",
                       "local_limbs_rank0": len(rows), "allgather_bytes_per_step": S * N * 8 if world > 1 else 0},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
