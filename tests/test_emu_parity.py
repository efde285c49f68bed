"""Kernel sources compiled for the host (tests/emu) vs the oracle, bit-exact, on small rings.

This exercises the real kernel code (tile NTT index maps, twiddle addressing, digit/extend/ModDown
tables, launch geometry, workspace layout) in a container without a GPU.  It is NOT a product path:
the product only ever loads the CUDA build (tiberate_fhe_b200/_native.py).  The GPU parity tests
proper are tests/test_gpu_parity.py (-m gpu).
"""

import pytest

import parity
from parity import Harness, Setup


@pytest.fixture(scope="module")
def h(emu_lib):
    return Harness(emu_lib, use_torch=False)


# (logN, num_scales, K): covers LA/LB = 4/4, 5/4, 5/5, 6/5, 4/8 and every tile-round pattern
CASES = [(8, 5, 2), (9, 4, 3), (10, 3, 1), (11, 6, 4), (12, 2, 2)]


@pytest.mark.parametrize("logN,ns,K", CASES)
def test_context_and_ntt(h, logN, ns, K):
    s = Setup.toy(h, logN, ns, K, seed=logN)
    try:
        parity.check_context(s)
        parity.check_ntt(s, level=0, with_special=True, batch=2)
        parity.check_ntt(s, level=min(1, ns), with_special=False, batch=1)
    finally:
        s.close()


@pytest.mark.parametrize("logN,ns,K", [(8, 5, 2), (10, 3, 1)])
def test_pointwise_and_he_ops(h, logN, ns, K):
    s = Setup.toy(h, logN, ns, K, seed=100 + logN)
    try:
        parity.check_pointwise(s, level=0, with_special=True)
        parity.check_pointwise(s, level=1, with_special=False)
        for level in range(0, ns + 1):
            parity.check_he_ops(s, level)
    finally:
        s.close()


@pytest.mark.parametrize("logN,ns,K", [(8, 5, 2), (9, 4, 3), (10, 3, 1)])
def test_engine_every_level(h, logN, ns, K):
    s = Setup.toy(h, logN, ns, K, seed=200 + logN, rot_deltas=(1, 3))
    try:
        # Host emulation budget: every level and every op in the default mode; the other modes (exact op
        # kernels, integer mod-q path) on the ops that contain transforms, at the first and the last level.
        # The GPU suite runs every mode at every level (tests/test_gpu_parity.py).
        for mode in (m for m in parity.ENGINE_MODES if m != (True, 3)):
            parity.set_mode(s, mode)
            if mode == parity.ENGINE_MODES[-1]:
                for level in range(0, ns + 1):
                    parity.check_engine(s, level)
            else:
                for level in (0, ns):
                    parity.check_engine(s, level, ops=("cc_mult", "rotate", "pc_mult"))
    finally:
        s.close()


def test_engine_logN12_compile_time_strides(h):
    """logN >= 12 selects the instantiations the presets run: LB = 8, compile-time strides, FP64-only row launches
    split from the integer rows, the fused key-switch core with its staged 4096-residue tiles."""
    s = Setup.toy(h, 12, 2, 2, seed=312, rot_deltas=(1, 2))
    try:
        parity.check_engine(s, 0, ops=("cc_mult", "rotate", "hoisted", "keyswitch", "pc_mult"))
        parity.check_engine(s, 1, batch=2, ops=("cc_mult", "rotate"))
    finally:
        s.close()


def test_engine_wide_scale_primes(h):
    """55-bit "scale" primes: every limb takes the reducing (non-small) butterfly policy, with
    multi-prime digit groups."""
    s = Setup.toy(h, 8, 4, 3, seed=11, scale_bits=55)
    try:
        for level in (0, 2, 4):
            parity.check_engine(s, level, ops=("keyswitch", "rotate", "cc_mult", "pc_mult"))
    finally:
        s.close()


def test_engine_primes_far_from_a_power_of_two(h):
    """Primes at 0.7 * 2^55 / 0.75 * 2^60: nothing may rely on q being just below a power of two."""
    s = Setup(h, 8, parity.far_primes(8, 4, 2), 2, seed=13)
    try:
        parity.check_ntt(s, level=0, with_special=True, batch=1)
        for level in (0, 3):
            parity.check_engine(s, level, ops=("keyswitch", "rotate", "cc_mult", "pc_mult"))
    finally:
        s.close()


def test_engine_batched_and_chunked(h):
    s = Setup.toy(h, 8, 4, 2, seed=7)
    try:
        s.ctx.set_chunk(2)  # batch 3 -> chunks of 2 + 1
        parity.check_engine(s, 0, batch=3)
        parity.check_engine(s, 2, batch=3, ops=("keyswitch", "rotate", "cc_mult", "pc_mult"))
    finally:
        s.close()


def test_bad_arguments_raise(h):
    import numpy as np

    from tiberate_fhe_b200 import Tb200Error
    from tiberate_fhe_b200.context import Tb200Context

    with pytest.raises(Tb200Error):
        Tb200Context(8, [17, 19, 23], 1, lib=h.lib)  # not NTT friendly
    s = Setup.toy(h, 8, 3, 1, seed=1)
    try:
        a = np.zeros((4, s.N), dtype=np.int64)
        with pytest.raises(Tb200Error):
            s.ctx.ntt(a, s.octx.P - 2, True)  # rows run past the last prime
        with pytest.raises(Tb200Error):
            s.ctx.ntt(np.zeros((4, s.N), dtype=np.int32), 0, True)  # wrong dtype
        with pytest.raises(Tb200Error):
            s.ctx.rescale(s.octx.num_ordinary - 1, a[:1], a[:1], a[:1], a[:1])  # no level left
        with pytest.raises(Tb200Error):
            s.ctx.rotate(0, 4, a, a, None, a.copy(), a.copy())  # even galois element
        # the engine layer derives every extent from `level`: operands with other row counts are refused
        # before the library sees their pointers
        with pytest.raises(Tb200Error, match="limb rows"):
            s.ctx.cc_addsub(1, False, a, a, a, a, a.copy(), a.copy())  # level 1 has 3 rows, not 4
        with pytest.raises(Tb200Error, match="limb rows"):
            s.ctx.cc_mult_relin(0, a, a, a[:3], a, s.evk_d, a[:3].copy(), a[:3].copy(), True)
        with pytest.raises(Tb200Error, match="batch"):
            b2 = np.zeros((2, 4, s.N), dtype=np.int64)
            s.ctx.cc_addsub(0, False, b2, b2, a, a, b2.copy(), b2.copy())
        with pytest.raises(Tb200Error, match="rows"):
            from tiberate_fhe_b200.context import KeySwitchKeyView

            short = KeySwitchKeyView([(np.zeros((2, s.N), dtype=np.int64), np.zeros((2, s.N), dtype=np.int64))], s.N)
            s.ctx.keyswitch(0, a, short, a.copy(), a.copy())
    finally:
        s.close()


@pytest.mark.parametrize("world", [2, 3])
def test_limb_sharded_keyswitch_matches_single_device(h, world):
    """BASELINE configs[3] logic on toy rings: every rank computes the ModUp digits of the digit groups
    it owns, the digit-state segments are all-gathered (emulated here by copies), each rank finishes
    the key switch for its own limbs; the assembled result equals the single-device oracle."""
    import numpy as np

    from oracle.context import OracleContext, toy_primes
    from oracle.engine import OracleEngine
    from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context

    logN, ns, K = 8, 7, 2
    q = toy_primes(logN, ns, K)
    octx = OracleContext(logN, q, K)
    eng = OracleEngine(octx)
    rng = np.random.default_rng(world)
    N = octx.N
    sk, _ = eng.gen_secret(rng)
    evk = eng.gen_evk(rng, sk)
    ctxs = [Tb200Context(logN, q, K, lib=h.lib, rank=r, world=world) for r in range(world)]
    try:
        owned = sorted(g for c in ctxs for g in c.local_prime_ids[: c.num_ordinary])
        assert owned == list(range(octx.num_ordinary)), "every ordinary prime has exactly one owner"
        for level in (0, 1, 3, 6):
            lp = octx.level_primes(level, False)
            a = eng.uniform(rng, lp)
            add = eng.uniform(rng, lp)
            want0, want1 = eng.create_switcher(a, evk, level)
            want_sw = eng.switch_key([add, a], evk, level)
            states, infos = [], []
            for c in ctxs:
                S, row0, seg, Lloc = c.ks_state_info(level)
                rows = [g - level for g in c.local_rows(level)]
                assert Lloc == len(rows)
                st = np.zeros((S, N), dtype=np.int64)
                a_loc = np.ascontiguousarray(a[rows]) if rows else np.zeros((1, N), dtype=np.int64)
                c.ks_digits(level, a_loc, st)
                states.append(st)
                infos.append((row0, seg, rows))
            for r, (row0, seg, _) in enumerate(infos):  # the all-gather
                for st in states:
                    st[row0:row0 + seg] = states[r][row0:row0 + seg]
            got0 = np.zeros_like(want0)
            got1 = np.zeros_like(want1)
            sw0 = np.zeros_like(want0)
            for c, st, (_, _, rows) in zip(ctxs, states, infos):
                if not rows:
                    continue
                ids = c.local_prime_ids
                key = KeySwitchKeyView([None if p is None else (np.ascontiguousarray(p[0][ids]),
                                                                np.ascontiguousarray(p[1][ids])) for p in evk], N)
                o0 = np.zeros((len(rows), N), dtype=np.int64)
                o1 = np.zeros_like(o0)
                c.ks_finish(level, st, key, o0, o1)
                got0[rows], got1[rows] = o0, o1
                c.ks_finish(level, st, key, o0, o1, add0=np.ascontiguousarray(add[rows]), tail=2)
                sw0[rows] = o0
            assert np.array_equal(got0, want0) and np.array_equal(got1, want1), f"level {level}"
            assert np.array_equal(sw0, want_sw[0]), f"switch_key tail, level {level}"
            # the same with the key sums of the special limbs sharded too (tb200_ks_core_sp / _ord / tb200_ks_moddown):
            # every rank -- also one without ordinary limbs left -- computes its share (world 3, K = 2: one rank has
            # none), the second all-gather is emulated by copies
            keys = []
            for c in ctxs:
                ids = c.local_prime_ids
                keys.append(KeySwitchKeyView([None if p is None else (np.ascontiguousarray(p[0][ids]),
                                                                      np.ascontiguousarray(p[1][ids])) for p in evk], N))
            sps = []
            for c, st, key in zip(ctxs, states, keys):
                rows_sp, seg_sp, s0, s1 = c.ks_sp_info()
                assert rows_sp == seg_sp * world and 0 <= s0 <= s1 <= K
                sp = np.zeros((rows_sp, 1, 2, N), dtype=np.int64)
                c.ks_modup(level, st, which=0 + 4)
                c.ks_core_sp(level, 1, key, sp)
                sps.append((sp, c.rank * seg_sp, seg_sp))
            for spr, r0, sg in sps:
                for sp, _, _ in sps:
                    if sp is not spr:
                        sp[r0:r0 + sg] = spr[r0:r0 + sg]
            got0, got1, sw0 = np.zeros_like(want0), np.zeros_like(want1), np.zeros_like(want0)
            for c, key, (_, _, rows), (sp, _, _) in zip(ctxs, keys, infos, sps):
                if not rows:
                    continue
                o0 = np.zeros((len(rows), N), dtype=np.int64)
                o1 = np.zeros_like(o0)
                c.ks_core_ord(level, 1, key, o0)
                sp2 = sp.copy()  # chain-backward rewrites the special limbs in place
                c.ks_moddown(level, sp, o0, o1)
                got0[rows], got1[rows] = o0, o1
                c.ks_moddown(level, sp2, o0, o1, add0=np.ascontiguousarray(add[rows]), tail=2)
                sw0[rows] = o0
            assert np.array_equal(got0, want0) and np.array_equal(got1, want1), f"sharded special, level {level}"
            assert np.array_equal(sw0, want_sw[0]), f"sharded special, switch_key tail, level {level}"
        # argument checks of the new entries: wrong sp shape; batch beyond the chunk
        from tiberate_fhe_b200 import Tb200Error

        c0 = ctxs[0]
        rows0 = len(c0.local_rows(0))
        with pytest.raises(Tb200Error, match="sp must be"):
            c0.ks_core_sp(0, 1, keys[0], np.zeros((1, 1, 2, N), dtype=np.int64)[:, :, :1])
        with pytest.raises(Tb200Error):
            c0.set_chunk(1)
            big = np.zeros((c0.ks_sp_info()[0], 2, 2, N), dtype=np.int64)
            c0.ks_core_sp(0, 2, keys[0], big)  # batch 2 > chunk 1
        _ = rows0
    finally:
        for c in ctxs:
            c.close()


def test_packed_wire_format_roundtrip(h):
    """tb200_pack41 / tb200_unpack41: 41 bits per residue (5 bytes + a bit plane), against a NumPy packing."""
    import numpy as np

    from tiberate_fhe_b200 import Tb200Error

    s = Setup.toy(h, 8, 3, 1, seed=4)
    try:
        ctx, N = s.ctx, s.N
        q = ctx.q
        nar = ctx.narrow_rows(0, ctx.P)
        assert nar == 3 and ctx.narrow_rows(1, 3) == 2 and ctx.packed_row_bytes == 5 * N + N // 8
        rng = np.random.default_rng(0)
        a = np.stack([np.stack([rng.integers(0, q[r], size=N, dtype=np.int64) for r in range(nar)]) for _ in range(2)])
        a[0, 0, 0], a[1, 2, N - 1] = q[0] - 1, q[2] - 1
        a[0, 1, 5] = (1 << 40) + 12345  # a residue that needs bit 40 (scale primes reach above 2^40)
        a[1, 0, 8] = 1 << 40
        packed = np.zeros((2, nar, ctx.packed_row_bytes), dtype=np.uint8)
        ctx.pack41(a, packed, 0)

        def ref_pack(row):
            lo = (row & ((1 << 40) - 1)).astype("<u8").view(np.uint8).reshape(N, 8)[:, :5].reshape(-1)
            hi = np.packbits(((row >> 40) & 1).astype(np.uint8), bitorder="little")
            return np.concatenate([lo, hi])

        ref = np.stack([np.stack([ref_pack(a[b, r]) for r in range(nar)]) for b in range(2)])
        assert np.array_equal(packed, ref)
        out = np.zeros_like(a)
        ctx.unpack41(packed, out, 0)
        assert np.array_equal(out, a)
        one = np.zeros((nar - 1, N), dtype=np.int64)  # unbatched, rows 1..: a row-offset view of the packed buffer
        ctx.unpack41(packed[1, 1:], one, 1)
        assert np.array_equal(one, a[1, 1:])
        with pytest.raises(Tb200Error, match="41 bits"):
            ctx.pack41(np.zeros((4, N), dtype=np.int64), np.zeros((4, ctx.packed_row_bytes), dtype=np.uint8), 0)  # row 3: 60-bit prime
    finally:
        s.close()
