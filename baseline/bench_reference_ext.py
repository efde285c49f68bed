"""Times the UNMODIFIED reference (its Python CkksEngine + its own CUDA extension rebuilt for
sm_100, pip-installed under baseline/_ref) on the same B200 and the same workload shape as bench.py:
logN16 preset, level 0, cc_mult (pre-rescale + relinearize), rotate_single, rescale, (i)NTT.

This is baseline (b) of BASELINE.json's north_star ("the reference's own CUDA extension built from
/root/reference on the same B200s"); a reported baseline, not the optimisation target.  Timing:
CUDA events around `iters` calls after `warmup` calls, torch.cuda.synchronize() at both ends (the
reference's own benchmark omits the synchronisation: extension/benchmarks/bench/single_cmult.py:83-88).
The reference has no batch dimension: ciphertexts are processed one per call.

  python baseline/bench_reference_ext.py [--iters 20] [--warmup 5] [--logN 16]   -> one JSON line
"""

from __future__ import annotations

import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--logN", type=int, default=16)
    args = ap.parse_args()
    import torch

    from baseline import ref_harness

    if not ref_harness.available() or not torch.cuda.is_available():
        print(json.dumps({"impl": "reference_cuda_ext", "unavailable": "baseline/_ref or CUDA device missing"}))
        return
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints its constant-pool layout
        ref_harness.load()
        from tiberate import CkksEngine, Preset

        engine = CkksEngine(getattr(Preset, f"logN{args.logN}"), devices=["cuda:0"])
        evk = engine.evk
        rotk = engine.rotk[1]
    N = engine.ckksCfg.N
    data = torch.randn(engine.num_slots, dtype=torch.float64)
    ct1 = engine.encodecrypt(data)
    ct2 = engine.encodecrypt(data * 0.5)
    L = ct1.data[0][0].shape[0]

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    res = {}
    res["cc_mult_relin_ms"] = timed(lambda: engine.cc_mult(ct1, ct2, evk))
    res["rotate_single_ms"] = timed(lambda: engine.rotate_single(ct1, rotk))
    res["rescale_ms"] = timed(lambda: engine.rescale(ct1))
    x = ct1.data[0][0].clone()

    def ntt_roundtrip():
        engine.nttCtx.enter_ntt_radix2([x], 0)
        engine.nttCtx.intt_radix2_exit_reduce([x], 0)

    res["ntt_fwd_plus_inv_ms"] = timed(ntt_roundtrip)
    # CSPRNG + self-contained engine path
    q_all = engine.nttCtx.q_prepack[-2][0][0]
    res["csprng_randint_allP_ms"] = timed(lambda: engine.rng.randint(q_all, repeats=engine.ckksCfg.num_special_primes))
    res["csprng_gaussian2_ms"] = timed(lambda: engine.rng.discrete_gaussian(repeats=2))
    res["encodecrypt_ms"] = timed(lambda: engine.encodecrypt(data))
    res["decryptcode_ms"] = timed(lambda: engine.decryptcode(ct1))
    # The same engine object on libtb200: operators only (the reference's Python still issues every op), then with
    # the hot methods bound to the fused entry points (tiberate_fhe_b200/backend.py) -- the drop-in numbers.
    dropin = {}
    try:
        from tiberate_fhe_b200 import backend

        for mode, fused in (("op_level", False), ("fused", True)):
            with contextlib.redirect_stdout(io.StringIO()):
                backend.install_as_tiberate_backend(engine, fused=fused)
            try:
                d = {"cc_mult_relin_ms": timed(lambda: engine.cc_mult(ct1, ct2, evk)),
                     "rotate_single_ms": timed(lambda: engine.rotate_single(ct1, rotk)),
                     "rescale_ms": timed(lambda: engine.rescale(ct1))}
                d["hmult_relin_ops_per_s"] = 1e3 / d["cc_mult_relin_ms"]
                d["speedup_vs_reference_ext"] = res["cc_mult_relin_ms"] / d["cc_mult_relin_ms"]
                dropin[mode] = d
            finally:
                backend.uninstall()
    except Exception as e:  # the baseline figures above stay valid
        dropin["error"] = repr(e)
    out = {
        "impl": "reference_cuda_ext", "logN": args.logN, "N": N, "limbs_level0": L, "dropin": dropin,
        "iters": args.iters, "warmup": args.warmup,
        "hmult_relin_ops_per_s": 1e3 / res["cc_mult_relin_ms"],
        "rotate_ops_per_s": 1e3 / res["rotate_single_ms"],
        "rescale_ops_per_s": 1e3 / res["rescale_ms"],
        "ntt_glimbs_per_s_fwd_inv_avg": 2 * L * N / (res["ntt_fwd_plus_inv_ms"] / 1e3) / 1e9,
        "csprng_randint_gsamples_per_s": len(engine.ckksCfg.q) * N / (res["csprng_randint_allP_ms"] / 1e3) / 1e9,
        "encodecrypt_ops_per_s": 1e3 / res["encodecrypt_ms"],
        "decryptcode_ops_per_s": 1e3 / res["decryptcode_ms"],
        **res,
        "note": "unmodified tiberate 0.9.11 engine + its CUDA extension (sm_100 build), one ciphertext per call, "
                "CUDA events with synchronisation",
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
