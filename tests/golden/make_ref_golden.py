"""Generates tests/golden/ref_logN*.json by running the UNMODIFIED reference (its Python engine and
its own CUDA extension, pip-installed under baseline/_ref and built for sm_100) on a B200.

  gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/golden'   # then copy the json files here

Inputs are seeded (tests/golden/golden_inputs.py); only hashes + a few words of every OUTPUT are
stored.  The checkers (tests/golden_check.py) recompute the same cases with the oracle (CPU) and
with libtb200 (GPU) and compare digests: this is what pins the oracle to the reference.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import golden_inputs as gi  # noqa: E402


def main(outdir: str, logNs=(14, 15, 16)):
    import torch

    from baseline import ref_harness

    ref_harness.load()
    from tiberate import CkksEngine, Preset
    from tiberate.libs.wrapper import he_ops, mont_ops, ntt2_ops  # noqa: F401
    from tiberate.typing import FLAGS, Ciphertext, CiphertextTriplet, KeySwitchKey, PublicKey, RotationKey

    os.makedirs(outdir, exist_ok=True)
    dev = "cuda:0"

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def H(t):
        return t.detach().cpu().numpy()

    for logN in logNs:
        engine = CkksEngine(getattr(Preset, f"logN{logN}"), devices=[dev])
        ctx = engine.nttCtx
        cfg = engine.ckksCfg
        q = [int(x) for x in cfg.q]
        N, K, P = cfg.N, cfg.num_special_primes, len(q)
        no = P - K
        out = {"logN": logN, "q": q, "K": K, "cases": {}}
        C = out["cases"]

        # ---- context constants -------------------------------------------------------------
        C["ctx/Rs"] = gi.record(H(ctx.Rs[0]))
        C["ctx/Ninv"] = gi.record(H(ctx.Ninv[0]))
        C["ctx/psi_expanded"] = gi.record(H(ctx.psi[0]))
        C["ctx/ipsi_expanded"] = gi.record(H(ctx.ipsi[0]))
        C["ctx/rescale_scales_l0"] = gi.record(H(engine.rescale_scales[0][0]))
        C["ctx/PiR_l0"] = {str(i): gi.record(H(engine.PiRs[0][i][0])) for i in range(K)}
        C["ctx/mont_PR"] = gi.record(H(engine.mont_PR[0]))
        out["digit_groups_l0"] = [list(p) for p in engine.rnsPart.destination_parts[0][0]]

        # ---- op level, all P rows (sp_prime_len = 0) -----------------------------------------
        x = gi.lazy_signed(100 + logN, q, N)
        y = gi.lazy_signed(200 + logN, q, N)
        tx, ty = T(x), T(y)
        for name in ("mont_mult", "mont_add", "mont_sub", "mont_add_reduce_2q", "mont_sub_reduce_2q"):
            C[f"op/{name}"] = gi.record(H(getattr(mont_ops, name)([tx], [ty], 0)[0]))
        for name in ("mont_enter_Rs", "mont_enter_Rs_scale", "mont_reduce", "reduce_2q", "make_signed", "make_unsigned"):
            t = tx.clone()
            getattr(mont_ops, name)([t], 0)
            C[f"op/{name}"] = gi.record(H(t))
        C["op/pc_add_fused"] = gi.record(H(he_ops.pc_add_fused([tx], [ty], 0)[0]))
        t = tx.clone()
        ntt2_ops.enter_ntt_radix2([t], ctx.even, ctx.odd, ctx.psi, 0)
        C["op/enter_ntt_radix2"] = gi.record(H(t))
        ent = t.clone()
        t = tx.clone()
        ntt2_ops.ntt_radix2([t], ctx.even, ctx.odd, ctx.psi, 0)
        C["op/ntt_radix2"] = gi.record(H(t))
        t = tx.clone()
        ntt2_ops.ntt_radix2([t], ctx.even, ctx.odd, ctx.psi, K)  # only the ordinary rows are transformed
        C["op/ntt_radix2_sp"] = gi.record(H(t))
        for name in ("intt_radix2", "intt_radix2_exit", "intt_radix2_exit_reduce", "intt_radix2_exit_reduce_signed"):
            t = ent.clone()
            getattr(ntt2_ops, name)([t], ctx.ieven, ctx.iodd, ctx.ipsi, 0)
            C[f"op/{name}"] = gi.record(H(t))
            t = ty.clone()  # lazy / negative residues straight into the inverse butterflies
            getattr(ntt2_ops, name)([t], ctx.ieven, ctx.iodd, ctx.ipsi, 0)
            C[f"op/{name}_lazy_input"] = gi.record(H(t))
        # ordinary rows only (sp_prime_len = K), level 2 tensor
        lv = 2
        xs = gi.uniform(300 + logN, q[lv:no], N)
        t = T(xs)
        ntt2_ops.enter_ntt_radix2([t], ctx.even, ctx.odd, ctx.psi, K)
        C["op/enter_ntt_radix2_level2"] = gi.record(H(t))
        ntt2_ops.intt_radix2_exit_reduce([t], ctx.ieven, ctx.iodd, ctx.ipsi, K)
        C["op/roundtrip_level2_is_identity"] = bool(np.array_equal(H(t), xs))

        # ---- engine level ------------------------------------------------------------------------
        ng = len(engine.rnsPart.destination_parts[0][0])
        flags = FLAGS.INCLUDE_SPECIAL | FLAGS.MONTGOMERY_STATE | FLAGS.NTT_STATE

        def make_key(seed, cls, **kw):
            parts = gi.ksk(seed, q, N, ng)
            data = [PublicKey(data=[[T(b)], [T(a)]], flags=flags, level=0) for b, a in parts]
            return cls(data=data, flags=flags, level=0, **kw)

        evk = make_key(1000 + logN, KeySwitchKey)
        rotk = make_key(2000 + logN, RotationKey, delta=gi.ROT_DELTA)

        def CT(polys, level):
            return Ciphertext(data=[[T(polys[0])], [T(polys[1])]], level=level)

        for level in gi.LEVEL_CASES[logN]:
            pr = q[level:no]
            c1 = gi.ciphertext(3000 + 10 * level + logN, pr, N)
            c2 = gi.ciphertext(4000 + 10 * level + logN, pr, N)
            tag = f"engine/l{level}"
            if level + 1 < engine.num_levels:
                r = engine.rescale(CT(c1, level))
                C[f"{tag}/rescale"] = [gi.record(H(r.data[0][0])), gi.record(H(r.data[1][0]))]
                r = engine.cc_mult(CT(c1, level), CT(c2, level), evk)
                C[f"{tag}/cc_mult_relin"] = [gi.record(H(r.data[0][0])), gi.record(H(r.data[1][0]))]
                r = engine.cc_mult(CT(c1, level), CT(c2, level), evk, post_relin=False)
                C[f"{tag}/cc_mult_triplet"] = [gi.record(H(d[0])) for d in r.data]
            r = engine.cc_mult(CT(c1, level), CT(c2, level), evk, pre_rescale=False)
            C[f"{tag}/cc_mult_relin_noprerescale"] = [gi.record(H(r.data[0][0])), gi.record(H(r.data[1][0]))]
            r = engine.rotate_single(CT(c1, level), rotk)
            C[f"{tag}/rotate_single"] = [gi.record(H(r.data[0][0])), gi.record(H(r.data[1][0]))]
            r = engine.switch_key(CT(c1, level), evk)
            C[f"{tag}/switch_key"] = [gi.record(H(r.data[0][0])), gi.record(H(r.data[1][0]))]
            r = engine.cc_add_double(CT(c1, level), CT(c2, level))
            C[f"{tag}/cc_add"] = [gi.record(H(r.data[0][0])), gi.record(H(r.data[1][0]))]
            r = engine.cc_sub_double(CT(c1, level), CT(c2, level)) if hasattr(engine, "cc_sub_double") else None
            if r is not None:
                C[f"{tag}/cc_sub"] = [gi.record(H(r.data[0][0])), gi.record(H(r.data[1][0]))]
            d0, d1 = engine.create_switcher([T(c1[1])], evk, level)
            C[f"{tag}/create_switcher"] = [gi.record(H(d0[0])), gi.record(H(d1[0]))]
        torch.cuda.synchronize()
        path = os.path.join(outdir, f"ref_logN{logN}.json")
        with open(path, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        print("wrote", path, len(C), "cases")
        del engine


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"),
         tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (14, 15, 16))
