#!/usr/bin/env python
"""bench_matvec.py -- BASELINE.json configs[4]: rotation-heavy encrypted matrix-vector product at
logN=16 (diagonal method: 64 x rotate_single + 64 x pc_mult + 63 x cc_add), exercising the Galois
automorphism + key switch throughput.

out = sum_{i<64} diag_i (.) rot(ct, i), with rot(ct, i) = rotate_single(rot(ct, i-1), rotk[1]).
A batch of B independent input ciphertexts shares the rotation key and the 64 plaintext diagonals
(cached NTT form, as Plaintext.cache in the reference); every step is one fused call over the batch.

  python bench_matvec.py [--batch 16] [--steps 3] [--warmup 1] [--level 0]      -> one JSON line
"""

from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def matvec(ctx, level, c0, c1, rotk, diags, galois, bufs):
    """One encrypted mat-vec over the batch. diags: [D, L, N] NTT-form plaintexts; returns (acc0, acc1)
    at level+1 (pc_mult rescales)."""
    r0, r1, t0, t1, p0, p1, acc0, acc1, s0, s1 = bufs
    cur0, cur1 = c0, c1
    for i in range(diags.shape[0]):
        if i > 0:
            nxt0, nxt1 = (r0, r1) if cur0 is not r0 else (t0, t1)
            ctx.rotate(level, galois, cur0, cur1, rotk, nxt0, nxt1)
            cur0, cur1 = nxt0, nxt1
        if i == 0:
            ctx.pc_mult(level, diags[i], cur0, cur1, acc0, acc1, post_rescale=True)
        else:
            ctx.pc_mult(level, diags[i], cur0, cur1, p0, p1, post_rescale=True)
            ctx.cc_addsub(level + 1, False, acc0, acc1, p0, p1, s0, s1)
            acc0, acc1, s0, s1 = s0, s1, acc0, acc1
    return acc0, acc1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--logN", type=int, default=16)
    args = ap.parse_args()
    import torch

    from tiberate_fhe_b200 import KeySwitchKeyView, Tb200Context, galois_element
    from tiberate_fhe_b200.presets import PRESETS

    dev = torch.device("cuda", 0)
    q, K = PRESETS[args.logN]["q"], PRESETS[args.logN]["K"]
    ctx = Tb200Context(args.logN, q, K)
    ctx.set_chunk(16)
    N, P, no = ctx.N, ctx.P, ctx.num_ordinary
    L = no - args.level
    B, D = args.batch, args.dim
    gen = torch.Generator(device=dev).manual_seed(0xB200)

    def uniform(shape, primes):
        t = torch.empty(*shape, dtype=torch.int64, device=dev)
        for i, qi in enumerate(primes):
            t[..., i, :].random_(0, int(qi), generator=gen)
        return t

    pr = q[args.level:no]
    c0, c1 = uniform((B, L, N), pr), uniform((B, L, N), pr)
    diags = uniform((D, L, N), pr)
    rotk = KeySwitchKeyView([(uniform((P, N), q), uniform((P, N), q)) for _ in range(ctx.num_groups0)], N)
    mk = lambda rows: torch.empty(B, rows, N, dtype=torch.int64, device=dev)  # noqa: E731
    bufs = [mk(L), mk(L), mk(L), mk(L), mk(L - 1), mk(L - 1), mk(L - 1), mk(L - 1), mk(L - 1), mk(L - 1)]
    g1 = galois_element(N, 1)
    for _ in range(args.warmup):
        matvec(ctx, args.level, c0, c1, rotk, diags, g1, bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        matvec(ctx, args.level, c0, c1, rotk, diags, g1, bufs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({
        "metric": f"encrypted {D}-diagonal mat-vec/s at logN={args.logN}", "value": B / (ms / 1e3), "unit": "matvec/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "dtype": "int64", "data": "synthetic",
        "rotations_per_s": B * (D - 1) / (ms / 1e3),
        "config": {"workload": f"logN{args.logN} preset, level {args.level}: {D - 1} rotate_single + {D} pc_mult(+rescale) + "
                               f"{D - 1} cc_add per mat-vec, batch {B} ciphertexts sharing key and diagonals"},
    }), flush=True)


if __name__ == "__main__":
    main()
