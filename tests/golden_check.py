"""Recomputes the cases of tests/golden/ref_logN*.json (outputs of the reference's own Python engine
+ CUDA extension on a B200, see tests/golden/make_ref_golden.py) with
  * the oracle (CPU)          -> pins the oracle to the reference,
  * libtb200 through the C ABI (GPU) -> pins the product to the reference,
and compares sha256 digests of every output tensor."""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import golden_inputs as gi  # noqa: E402

from oracle import _c, call  # noqa: E402
from oracle.context import OracleContext  # noqa: E402
from oracle.engine import OracleEngine  # noqa: E402


def _cmp(name, got, want):
    if isinstance(want, list):
        assert len(got) == len(want), name
        for i, (g, w) in enumerate(zip(got, want)):
            _cmp(f"{name}[{i}]", g, w)
        return
    g = np.ascontiguousarray(got, dtype=np.int64)
    assert list(g.shape) == want["shape"], f"{name}: shape {g.shape} vs reference {want['shape']}"
    if gi.digest(g) != want["sha256"]:
        raise AssertionError(
            f"{name}: digest differs from the reference extension; head {g.reshape(-1)[:4].tolist()} vs "
            f"{want['head']}, tail {g.reshape(-1)[-4:].tolist()} vs {want['tail']}")


def expanded_tables(octx: OracleContext):
    """the expanded [P, logN, N/2] twiddle tensors from the compact tables."""
    logN, N = octx.logN, octx.N
    j = np.arange(N // 2)
    psi = np.empty((octx.P, logN, N // 2), dtype=np.int64)
    ipsi = np.empty_like(psi)
    for s in range(logN):
        psi[:, s, :] = octx.psi[:, (1 << s) + j // (N >> (s + 1))]
        ipsi[:, s, :] = octx.ipsi[:, (1 << (logN - 1 - s)) + (j >> s)]
    return psi, ipsi


# ------------------------------------------------------------------------------------------------
def oracle_cases(meta):
    logN, q, K = meta["logN"], meta["q"], meta["K"]
    octx = OracleContext(logN, q, K)
    eng = OracleEngine(octx)
    N, P, no = octx.N, octx.P, octx.num_ordinary
    allp = list(range(P))
    qa, ka = octx.rows(allp)
    out = {}
    out["ctx/Rs"] = octx.Rsa
    out["ctx/Ninv"] = octx.Ninva
    psi, ipsi = expanded_tables(octx)
    out["ctx/psi_expanded"], out["ctx/ipsi_expanded"] = psi, ipsi
    out["ctx/rescale_scales_l0"] = octx.rescale_scales[0]
    out["ctx/mont_PR"] = octx.mont_PR
    x = gi.lazy_signed(100 + logN, q, N)
    y = gi.lazy_signed(200 + logN, q, N)

    def two(name, *extra):
        r = np.empty_like(x)
        call(name, r, _c(x), _c(y), P, N, *extra)
        return r

    def one(name, *extra):
        r = _c(x).copy()
        call(name, r, *extra)
        return r

    out["op/mont_mult"] = two("orc_mont_mult", qa, ka)
    out["op/mont_add"] = two("orc_mont_add", qa)
    out["op/mont_sub"] = two("orc_mont_sub", qa)
    out["op/mont_add_reduce_2q"] = two("orc_mont_add_reduce_2q", qa)
    out["op/mont_sub_reduce_2q"] = two("orc_mont_sub_reduce_2q", qa)
    out["op/mont_enter_Rs"] = one("orc_mont_enter_scalar", octx.Rsa, P, N, qa, ka)
    out["op/mont_enter_Rs_scale"] = one("orc_mont_enter_scalar", _c(octx.Rs_scale), P, N, qa, ka)
    out["op/mont_reduce"] = one("orc_mont_reduce", P, N, qa, ka)
    out["op/reduce_2q"] = one("orc_reduce_2q", P, N, qa)
    out["op/make_signed"] = one("orc_make_signed", P, N, qa)
    out["op/make_unsigned"] = one("orc_make_unsigned", P, N, qa)
    r = np.empty_like(x)
    call("orc_pc_add_fused", r, _c(x), _c(y), P, N, qa, ka, octx.Rsa)
    out["op/pc_add_fused"] = r
    ent = eng.enter_ntt(x, allp)
    out["op/enter_ntt_radix2"] = ent
    out["op/ntt_radix2"] = eng.ntt(x, allp)
    t = x.copy()
    t[:no] = eng.ntt(x[:no], allp[:no])
    out["op/ntt_radix2_sp"] = t
    for mode, name in enumerate(("intt_radix2", "intt_radix2_exit", "intt_radix2_exit_reduce",
                                 "intt_radix2_exit_reduce_signed")):
        out[f"op/{name}"] = eng.intt(ent, allp, mode)
        out[f"op/{name}_lazy_input"] = eng.intt(y, allp, mode)
    lv = 2
    xs = gi.uniform(300 + logN, q[lv:no], N)
    out["op/enter_ntt_radix2_level2"] = eng.enter_ntt(xs, list(range(lv, no)))
    ng = octx.part.num_partitions + 1
    evk = gi.ksk(1000 + logN, q, N, ng)
    rotk = gi.ksk(2000 + logN, q, N, ng)
    for level in gi.LEVEL_CASES[logN]:
        pr = q[level:no]
        c1 = gi.ciphertext(3000 + 10 * level + logN, pr, N)
        c2 = gi.ciphertext(4000 + 10 * level + logN, pr, N)
        tag = f"engine/l{level}"
        if level + 1 < octx.num_levels:
            out[f"{tag}/rescale"] = eng.rescale(c1, level)
            out[f"{tag}/cc_mult_relin"] = eng.cc_mult(c1, c2, evk, level)[0]
            out[f"{tag}/cc_mult_triplet"] = eng.cc_mult(c1, c2, evk, level, post_relin=False)[0]
        out[f"{tag}/cc_mult_relin_noprerescale"] = eng.cc_mult(c1, c2, evk, level, pre_rescale=False)[0]
        out[f"{tag}/rotate_single"] = eng.rotate_single(c1, rotk, gi.ROT_DELTA, level)
        out[f"{tag}/switch_key"] = eng.switch_key(c1, evk, level)
        out[f"{tag}/cc_add"] = eng.cc_add(c1, c2, level)
        out[f"{tag}/cc_sub"] = eng.cc_sub(c1, c2, level)
        out[f"{tag}/create_switcher"] = list(eng.create_switcher(c1[1], evk, level))
    return out, octx


def lib_cases(h, meta):
    """The same cases through libtb200 (h: parity.Harness)."""
    from tiberate_fhe_b200.context import Tb200Context, galois_element

    logN, q, K = meta["logN"], meta["q"], meta["K"]
    ctx = Tb200Context(logN, q, K, device=h.device, lib=h.lib)
    N, P, no = ctx.N, ctx.P, ctx.num_ordinary
    out = {}
    pc = ctx.prime_consts()
    out["ctx/Rs"], out["ctx/Ninv"] = pc[:, 4].copy(), pc[:, 6].copy()
    j = np.arange(N // 2)
    psi = np.empty((P, logN, N // 2), dtype=np.int64)
    ipsi = np.empty_like(psi)
    for g in range(P):
        f, b = ctx.twiddles(False, g), ctx.twiddles(True, g)
        for s in range(logN):
            psi[g, s] = f[(1 << s) + j // (N >> (s + 1))]
            ipsi[g, s] = b[(1 << (logN - 1 - s)) + (j >> s)]
    out["ctx/psi_expanded"], out["ctx/ipsi_expanded"] = psi, ipsi
    x = gi.lazy_signed(100 + logN, q, N)
    y = gi.lazy_signed(200 + logN, q, N)
    dx, dy = h.dev(x), h.dev(y)
    for op, name in ((0, "mont_mult"), (1, "mont_add"), (2, "mont_sub"), (3, "mont_add_reduce_2q"),
                     (4, "mont_sub_reduce_2q"), (13, "pc_add_fused")):
        o = h.zeros(P, N)
        ctx.pointwise(op, dx, dy, o, 0)
        out[f"op/{name}"] = h.host(o)
    for op, name in ((6, "mont_enter_Rs"), (7, "mont_enter_Rs_scale"), (8, "mont_reduce"), (9, "reduce_2q"),
                     (10, "make_signed"), (11, "make_unsigned")):
        t = h.dev(x)
        ctx.pointwise(op, t, None, None, 0)
        out[f"op/{name}"] = h.host(t)
    t = h.dev(x)
    ctx.ntt(t, 0, True)
    ent = h.host(t)
    out["op/enter_ntt_radix2"] = ent
    t = h.dev(x)
    ctx.ntt(t, 0, False)
    out["op/ntt_radix2"] = h.host(t)
    t = h.dev(x)
    ctx.ntt(t, 0, False, rows=no)
    out["op/ntt_radix2_sp"] = h.host(t)
    for mode, name in enumerate(("intt_radix2", "intt_radix2_exit", "intt_radix2_exit_reduce",
                                 "intt_radix2_exit_reduce_signed")):
        t = h.dev(ent)
        ctx.intt(t, 0, mode)
        out[f"op/{name}"] = h.host(t)
        t = h.dev(y)
        ctx.intt(t, 0, mode)
        out[f"op/{name}_lazy_input"] = h.host(t)
    lv = 2
    t = h.dev(gi.uniform(300 + logN, q[lv:no], N))
    ctx.ntt(t, lv, True)
    out["op/enter_ntt_radix2_level2"] = h.host(t)
    ng = ctx.num_groups0
    evk = h.key(gi.ksk(1000 + logN, q, N, ng), N)
    rotk = h.key(gi.ksk(2000 + logN, q, N, ng), N)
    for level in gi.LEVEL_CASES[logN]:
        pr = q[level:no]
        L = len(pr)
        c1 = [h.dev(t_) for t_ in gi.ciphertext(3000 + 10 * level + logN, pr, N)]
        c2 = [h.dev(t_) for t_ in gi.ciphertext(4000 + 10 * level + logN, pr, N)]
        tag = f"engine/l{level}"

        def pair(rows):
            return h.zeros(rows, N), h.zeros(rows, N)

        if level + 1 < ctx.num_scales:
            o0, o1 = pair(L - 1)
            ctx.rescale(level, c1[0], c1[1], o0, o1)
            out[f"{tag}/rescale"] = [h.host(o0), h.host(o1)]
            o0, o1 = pair(L - 1)
            ctx.cc_mult_relin(level, c1[0], c1[1], c2[0], c2[1], evk, o0, o1, True)
            out[f"{tag}/cc_mult_relin"] = [h.host(o0), h.host(o1)]
            d = [h.zeros(L - 1, N) for _ in range(3)]
            ctx.cc_mult_triplet(level, c1[0], c1[1], c2[0], c2[1], d[0], d[1], d[2], True)
            out[f"{tag}/cc_mult_triplet"] = [h.host(t_) for t_ in d]
        o0, o1 = pair(L)
        ctx.cc_mult_relin(level, c1[0], c1[1], c2[0], c2[1], evk, o0, o1, False)
        out[f"{tag}/cc_mult_relin_noprerescale"] = [h.host(o0), h.host(o1)]
        o0, o1 = pair(L)
        ctx.rotate(level, galois_element(N, gi.ROT_DELTA), c1[0], c1[1], rotk, o0, o1)
        out[f"{tag}/rotate_single"] = [h.host(o0), h.host(o1)]
        o0, o1 = pair(L)
        ctx.switch_key(level, c1[0], c1[1], evk, o0, o1)
        out[f"{tag}/switch_key"] = [h.host(o0), h.host(o1)]
        o0, o1 = pair(L)
        ctx.cc_addsub(level, False, c1[0], c1[1], c2[0], c2[1], o0, o1)
        out[f"{tag}/cc_add"] = [h.host(o0), h.host(o1)]
        o0, o1 = pair(L)
        ctx.cc_addsub(level, True, c1[0], c1[1], c2[0], c2[1], o0, o1)
        out[f"{tag}/cc_sub"] = [h.host(o0), h.host(o1)]
        o0, o1 = pair(L)
        ctx.keyswitch(level, c1[1], evk, o0, o1)
        out[f"{tag}/create_switcher"] = [h.host(o0), h.host(o1)]
    ctx.close()
    return out


def compare(meta, computed, who: str):
    checked = 0
    for name, want in meta["cases"].items():
        if name not in computed:
            continue
        if isinstance(want, bool):
            assert want, f"reference self-check {name} failed"
            continue
        if isinstance(want, dict) and "sha256" not in want:  # nested dict (ctx/PiR_l0): skip, covered by moddown
            continue
        _cmp(f"{who}:{name}", computed[name], want)
        checked += 1
    assert checked > 0
    return checked


def load(path):
    with open(path) as f:
        return json.load(f)


def check_file(h, path):
    meta = load(path)
    return compare(meta, lib_cases(h, meta), "libtb200")


def check_file_oracle(path):
    meta = load(path)
    from tiberate_fhe_b200.presets import PRESETS

    assert meta["q"] == PRESETS[meta["logN"]]["q"] and meta["K"] == PRESETS[meta["logN"]]["K"], \
        "prime chain produced by the reference differs from the pinned preset"
    cases, octx = oracle_cases(meta)
    groups = [pr for _, pr in octx.part.level_groups(0)]
    assert groups == meta["digit_groups_l0"], "digit groups differ from the reference's RnsPartition"
    return compare(meta, cases, "oracle")
