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


def matvec_hoisted(ctx, level, c0, c1, rotks, galois, diags, bufs):
    """The same product with every rotation taken from the INPUT ciphertext (rot(ct, i) with key rotk[i]), so the
    D - 1 rotations share one ModUp (tb200_rotate_hoisted): digits, extension and forward transform once."""
    rot0, rot1, p0, p1, acc0, acc1, s0, s1 = bufs
    ctx.rotate_hoisted(level, galois, c0, c1, rotks, rot0, rot1)
    ctx.pc_mult(level, diags[0], c0, c1, acc0, acc1, post_rescale=True)
    for i in range(1, diags.shape[0]):
        ctx.pc_mult(level, diags[i], rot0[i - 1], rot1[i - 1], p0, p1, post_rescale=True)
        ctx.cc_addsub(level + 1, False, acc0, acc1, p0, p1, s0, s1)
        acc0, acc1, s0, s1 = s0, s1, acc0, acc1
    return acc0, acc1


def matvec_bsgs(ctx, level, c0, c1, baby_keys, baby_gal, giant_keys, giant_gal, diags, bufs, n1, n2):
    """Baby-step / giant-step: out = sum_j rot_{j n1}( sum_i diag'_{ij} (.) rot_i(ct) ): n1 - 1 hoisted baby
    rotations of the input, n1 n2 pc_mult, n2 - 1 giant rotations (one per partial sum, at level + 1)."""
    rot0, rot1, p0, p1, in0, in1, s0, s1, g0, g1, acc0, acc1 = bufs
    ctx.rotate_hoisted(level, baby_gal, c0, c1, baby_keys, rot0, rot1)
    for j in range(n2):
        for i in range(n1):
            src0, src1 = (c0, c1) if i == 0 else (rot0[i - 1], rot1[i - 1])
            if i == 0:
                ctx.pc_mult(level, diags[j * n1], src0, src1, in0, in1, post_rescale=True)
            else:
                ctx.pc_mult(level, diags[j * n1 + i], src0, src1, p0, p1, post_rescale=True)
                ctx.cc_addsub(level + 1, False, in0, in1, p0, p1, s0, s1)
                in0, in1, s0, s1 = s0, s1, in0, in1
        if j == 0:
            acc0.copy_(in0)
            acc1.copy_(in1)
        else:
            ctx.rotate(level + 1, giant_gal[j - 1], in0, in1, giant_keys[j - 1], g0, g1)
            ctx.cc_addsub(level + 1, False, acc0, acc1, g0, g1, s0, s1)
            acc0, acc1, s0, s1 = s0, s1, acc0, acc1
    return acc0, acc1


def run(mode="hoisted", batch=16, steps=3, warmup=1, level=0, dim=64, logN=16, chunk=16):
    """One measurement; returns the JSON-able result."""
    import torch

    from tiberate_fhe_b200 import KeySwitchKeyView, Tb200Context, galois_element
    from tiberate_fhe_b200.presets import PRESETS

    dev = torch.device("cuda", torch.cuda.current_device())
    q, K = PRESETS[logN]["q"], PRESETS[logN]["K"]
    ctx = Tb200Context(logN, q, K, device=dev.index)
    ctx.set_chunk(chunk)
    N, P, no = ctx.N, ctx.P, ctx.num_ordinary
    L = no - level
    B, D = batch, dim
    gen = torch.Generator(device=dev).manual_seed(0xB200)

    def uniform(shape, primes):
        t = torch.empty(*shape, dtype=torch.int64, device=dev)
        for i, qi in enumerate(primes):
            t[..., i, :].random_(0, int(qi), generator=gen)
        return t

    def key():
        return KeySwitchKeyView([(uniform((P, N), q), uniform((P, N), q)) for _ in range(ctx.num_groups0)], N)

    pr = q[level:no]
    c0, c1 = uniform((B, L, N), pr), uniform((B, L, N), pr)
    diags = uniform((D, L, N), pr)
    mk = lambda *lead: torch.empty(*lead, N, dtype=torch.int64, device=dev)  # noqa: E731
    if mode == "chain":
        rotk = key()
        bufs = [mk(B, L), mk(B, L), mk(B, L), mk(B, L)] + [mk(B, L - 1) for _ in range(6)]
        g1 = galois_element(N, 1)
        fn = lambda: matvec(ctx, level, c0, c1, rotk, diags, g1, bufs)  # noqa: E731
        nrot, nkeys = D - 1, 1
    elif mode == "hoisted":
        keys = [key() for _ in range(D - 1)]
        gal = [galois_element(N, i) for i in range(1, D)]
        bufs = [mk(D - 1, B, L), mk(D - 1, B, L)] + [mk(B, L - 1) for _ in range(6)]
        fn = lambda: matvec_hoisted(ctx, level, c0, c1, keys, gal, diags, bufs)  # noqa: E731
        nrot, nkeys = D - 1, D - 1
    else:
        n1 = 1
        while n1 * n1 < D:
            n1 *= 2
        n2 = D // n1
        bk = [key() for _ in range(n1 - 1)]
        gk = [key() for _ in range(n2 - 1)]
        bg = [galois_element(N, i) for i in range(1, n1)]
        gg = [galois_element(N, j * n1) for j in range(1, n2)]
        bufs = [mk(n1 - 1, B, L), mk(n1 - 1, B, L)] + [mk(B, L - 1) for _ in range(10)]
        fn = lambda: matvec_bsgs(ctx, level, c0, c1, bk, bg, gk, gg, diags, bufs, n1, n2)  # noqa: E731
        nrot, nkeys = n1 - 1 + n2 - 1, n1 + n2 - 2
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.close()
    return {
        "metric": f"encrypted {D}-diagonal mat-vec/s at logN={logN}", "value": B / (ms / 1e3), "unit": "matvec/s",
        "mode": mode, "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms, "dtype": "int64",
        "data": "synthetic", "rotations_per_matvec": nrot, "rotation_keys": nkeys,
        "rotations_per_s": B * nrot / (ms / 1e3),
        "config": {"workload": f"logN{logN} preset, level {level}: {nrot} rotations + {D} pc_mult(+rescale) + cc_add per "
                               f"mat-vec, batch {B} ciphertexts sharing keys and diagonals",
                   "modes": "chain: rot^i by repeated rotate_single(rotk[1]); hoisted: rot_i of the input with key rotk[i], "
                            "one ModUp for all (tb200_rotate_hoisted); bsgs: hoisted baby steps + per-sum giant steps"},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="all", choices=["chain", "hoisted", "bsgs", "all"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--logN", type=int, default=16)
    args = ap.parse_args()
    for mode in (("chain", "hoisted", "bsgs") if args.mode == "all" else (args.mode,)):
        print(json.dumps(run(mode, args.batch, args.steps, args.warmup, args.level, args.dim, args.logN)), flush=True)


if __name__ == "__main__":
    main()
