"""GPU parity at the shapes BASELINE.json's configs name and bench.py times ("parity gaps"):

  * config 3 (headline): logN16 against the oracle in every internal mode, and one batched call that crosses
    a chunk boundary at the benchmark's chunk size (33 ciphertext pairs, chunk 32);
  * config 4: the limb-sharded key switch at logN17 (79 primes, 13 digit groups) against the ORACLE, with the
    ranks emulated on one GPU (sharded contexts side by side, the all-gather done by copies) so that a
    1-GPU box still produces correctness evidence for the sharded code path (the NCCL run itself is
    tests/test_gpu_sharded.py, which needs two GPUs);
  * config 5: an 8-diagonal encrypted matrix-vector product at logN14 with real keys: the integer result
    equals the oracle's composition of rotate_single / pc_mult / cc_add, and decrypts to the plaintext
    product within the scheme's error.
"""

import numpy as np
import pytest

import parity
from parity import Harness, Setup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h():
    from tiberate_fhe_b200 import get_lib

    return Harness(get_lib(), use_torch=True)


def _preset(h, logN, **kw):
    from oracle.context import PRESETS

    return Setup(h, logN, PRESETS[logN]["q"], PRESETS[logN]["K"], seed=1000 + logN, **kw)


def test_logN16_every_internal_mode(h):
    s = _preset(h, 16)
    try:
        for mode in parity.ENGINE_MODES:
            parity.set_mode(s, mode)
            parity.check_engine(s, 0, ops=("cc_mult", "rotate"))
            parity.check_engine(s, 31, ops=("keyswitch", "switch_key", "triplet"))
    finally:
        s.close()


def test_logN16_batch_crosses_a_chunk_boundary(h):
    """33 ciphertext pairs at the benchmark's chunk of 32: chunks of 32 + 1, every entry against the oracle."""
    s = _preset(h, 16)
    try:
        s.ctx.set_chunk(32)
        B, level = 33, 0
        o, eng, N = s.octx, s.eng, s.N
        L = o.num_ordinary - level
        ct1, ct2 = s.ct(level, B), s.ct(level, B)
        d1 = [h.dev(x) for x in ct1]
        d2 = [h.dev(x) for x in ct2]
        o0, o1 = h.zeros(B, L - 1, N), h.zeros(B, L - 1, N)
        s.ctx.cc_mult_relin(level, d1[0], d1[1], d2[0], d2[1], s.evk_d, o0, o1, True)
        g0, g1 = h.host(o0), h.host(o1)
        for b in range(B):
            w = eng.cc_mult([ct1[0][b], ct1[1][b]], [ct2[0][b], ct2[1][b]], s.evk, level, pre_rescale=True)[0]
            assert np.array_equal(g0[b], w[0]) and np.array_equal(g1[b], w[1]), f"batch entry {b}"
        r0, r1 = h.zeros(B, L, N), h.zeros(B, L, N)
        from tiberate_fhe_b200.context import galois_element

        s.ctx.rotate(level, galois_element(N, 1), d1[0], d1[1], s.rotk_d[1], r0, r1)
        g0, g1 = h.host(r0), h.host(r1)
        for b in (0, 31, 32):
            w = eng.rotate_single([ct1[0][b], ct1[1][b]], s.rotk[1], 1, level)
            assert np.array_equal(g0[b], w[0]) and np.array_equal(g1[b], w[1]), f"rotate, batch entry {b}"
    finally:
        s.close()


@pytest.mark.parametrize("world", [2, 4])
def test_logN17_limb_sharded_keyswitch_vs_oracle(h, world):
    import sys, os

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import golden_inputs as gi
    import torch

    from oracle.context import OracleContext
    from oracle.engine import OracleEngine
    from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context
    from tiberate_fhe_b200.presets import PRESETS

    q, K = PRESETS[17]["q"], PRESETS[17]["K"]
    octx = OracleContext(17, q, K)
    eng = OracleEngine(octx)
    N, no = octx.N, octx.num_ordinary
    ng = octx.part.num_partitions + 1
    evk = gi.ksk(1700, q, N, ng)
    ctxs = [Tb200Context(17, q, K, device=h.device, lib=h.lib, rank=r, world=world) for r in range(world)]
    try:
        keys = []
        for c in ctxs:
            ids = c.local_prime_ids
            keys.append(KeySwitchKeyView([(h.dev(p[0][ids]), h.dev(p[1][ids])) for p in evk], N))
        for level in (40, 66):
            lp = q[level:no]
            a = gi.uniform(1701 + level, lp, N)
            add = gi.uniform(1702 + level, lp, N)
            want0, want1 = eng.create_switcher(a, evk, level)
            want_sw = eng.switch_key([add, a], evk, level)
            states, infos = [], []
            for c in ctxs:
                S, row0, seg, Lloc = c.ks_state_info(level)
                rows = [g - level for g in c.local_rows(level)]
                assert Lloc == len(rows)
                st = h.zeros(S, N)
                if rows:
                    c.ks_digits(level, h.dev(a[rows]), st)
                states.append(st)
                infos.append((row0, seg, rows))
            torch.cuda.synchronize()
            for r, (row0, seg, _) in enumerate(infos):  # what the NCCL all-gather does (dist.py)
                for st in states:
                    if st is not states[r]:
                        st[row0:row0 + seg].copy_(states[r][row0:row0 + seg])
            got0, got1, sw0 = np.zeros_like(want0), np.zeros_like(want1), np.zeros_like(want0)
            for c, key, st, (_, _, rows) in zip(ctxs, keys, states, infos):
                if not rows:
                    continue
                o0, o1 = h.zeros(len(rows), N), h.zeros(len(rows), N)
                c.ks_finish(level, st, key, o0, o1)
                got0[rows], got1[rows] = h.host(o0), h.host(o1)
                c.ks_finish(level, st, key, o0, o1, add0=h.dev(add[rows]), tail=2)
                sw0[rows] = h.host(o0)
            assert np.array_equal(got0, want0) and np.array_equal(got1, want1), f"world {world} level {level}"
            assert np.array_equal(sw0, want_sw[0]), f"switch_key tail, world {world} level {level}"
            # the same with the key sums of the special limbs sharded as well (tb200_ks_core_sp): every rank, also one
            # that has run out of ordinary limbs, computes its share; the second all-gather is emulated by copies
            sps = []
            for c, key, st in zip(ctxs, keys, states):
                rows_sp, seg_sp, s0, s1 = c.ks_sp_info()
                sp = torch.zeros(rows_sp, 1, 2, N, dtype=torch.int64, device=st.device)
                c.ks_modup(level, st, which=0 + 4)
                c.ks_core_sp(level, 1, key, sp)
                sps.append((sp, c.rank * seg_sp, seg_sp))
            torch.cuda.synchronize()
            for r, (spr, r0, sg) in enumerate(sps):
                for sp, _, _ in sps:
                    if sp is not spr:
                        sp[r0:r0 + sg].copy_(spr[r0:r0 + sg])
            got0, got1, sw0 = np.zeros_like(want0), np.zeros_like(want1), np.zeros_like(want0)
            for c, key, st, (_, _, rows), (sp, _, _) in zip(ctxs, keys, states, infos, sps):
                if not rows:
                    continue
                o0, o1 = h.zeros(len(rows), N), h.zeros(len(rows), N)
                c.ks_core_ord(level, 1, key, o0)
                sp2 = sp.clone()  # chain-backward rewrites the special limbs in place
                c.ks_moddown(level, sp, o0, o1)
                got0[rows], got1[rows] = h.host(o0), h.host(o1)
                c.ks_moddown(level, sp2, o0, o1, add0=h.dev(add[rows]), tail=2)
                sw0[rows] = h.host(o0)
            assert np.array_equal(got0, want0) and np.array_equal(got1, want1), f"sharded special, world {world} level {level}"
            assert np.array_equal(sw0, want_sw[0]), f"sharded special, switch_key tail, world {world} level {level}"
    finally:
        for c in ctxs:
            c.close()


def _ksk_numpy(key):
    """tiberate_fhe_b200 KeySwitchKey -> the oracle's list of (b, a) arrays per digit group."""
    return [(p.data[0][0].cpu().numpy(), p.data[1][0].cpu().numpy()) for p in key.data]


def test_config5_matvec_logN14_oracle_composition_and_decrypt():
    import torch

    import tiberate_fhe_b200 as tb
    from oracle.context import OracleContext
    from oracle.engine import OracleEngine
    from tiberate_fhe_b200.typing import Plaintext

    D = 8
    eng = tb.CkksEngine(14, devices=["cuda:0"], seed=list(range(1, 9)), nonce=[3, 4])
    octx = OracleContext(14, list(eng.ctx.q), eng.ctx.K)
    orc = OracleEngine(octx)
    gen = torch.Generator().manual_seed(5)
    n = eng.num_slots
    v = torch.randn(n, generator=gen, dtype=torch.float64)
    diags = [torch.randn(n, generator=gen, dtype=torch.float64) for _ in range(D)]
    ct = eng.encodecrypt(v)
    rotk1 = eng.rotk[1]
    pts = [Plaintext(d) for d in diags]
    # the engine's mat-vec: out = sum_i diag_i (.) rot^i(ct)
    cur, acc = ct, None
    for i in range(D):
        if i > 0:
            cur = eng.rotate_single(cur, rotk1)
        term = eng.pc_mult(pts[i], cur)
        acc = term if acc is None else eng.cc_add(acc, term)
    # 1. bit-exact against the oracle's composition of the same operators on the same keys and plaintexts
    level = ct.level
    rk = _ksk_numpy(rotk1)
    ocur = [ct.data[0][0].cpu().numpy(), ct.data[1][0].cpu().numpy()]
    oacc = None
    for i in range(D):
        if i > 0:
            ocur = orc.rotate_single(ocur, rk, 1, level)
        pt_ntt = pts[i].cache[level]["pc_mult"][0].cpu().numpy()
        term, lvl1 = orc.pc_mult(pt_ntt, ocur, level, post_rescale=True)
        oacc = term if oacc is None else orc.cc_add(oacc, term, lvl1)
    assert acc.level == lvl1
    assert np.array_equal(acc.data[0][0].cpu().numpy(), oacc[0]), "mat-vec c0 differs from the oracle composition"
    assert np.array_equal(acc.data[1][0].cpu().numpy(), oacc[1]), "mat-vec c1 differs from the oracle composition"
    # 1b. the same product with hoisted rotations (rot_i of the input, key rotk[i], one ModUp for all): every rotation
    # equals the oracle's restatement of the hoisted algorithm bit for bit, and the sum decrypts to the same product
    hoisted = eng.rotate_hoisted(ct, range(1, D))
    ohoist = orc.rotate_hoisted([ct.data[0][0].cpu().numpy(), ct.data[1][0].cpu().numpy()],
                                {d: _ksk_numpy(eng.rotk[d]) for d in range(1, D)}, list(range(1, D)), level)
    for r, (got, want) in enumerate(zip(hoisted, ohoist)):
        assert np.array_equal(got.data[0][0].cpu().numpy(), want[0]), f"hoisted rotation {r + 1} c0"
        assert np.array_equal(got.data[1][0].cpu().numpy(), want[1]), f"hoisted rotation {r + 1} c1"
    acc_h = eng.pc_mult(pts[0], ct)
    for i in range(1, D):
        acc_h = eng.cc_add(acc_h, eng.pc_mult(pts[i], hoisted[i - 1]))
    dec_h = torch.as_tensor(np.asarray(eng.decryptcode(acc_h, is_real=True)), dtype=torch.float64)[:n]
    # 2. decrypts to the plaintext product (either rotation direction convention, fixed over the sum)
    dec = torch.as_tensor(np.asarray(eng.decryptcode(acc, is_real=True)), dtype=torch.float64)[:n]
    errs = []
    for sgn in (-1, 1):
        want = sum(diags[i] * torch.roll(v, sgn * i) for i in range(D))
        errs.append((dec - want).abs().max().item())
    tol = 1e-4 * max(1.0, float(sum(d.abs().max() for d in diags) * v.abs().max()))
    assert min(errs) < tol, errs
    assert (dec_h - dec).abs().max().item() < tol, "hoisted mat-vec decrypts to something else than the chained one"
