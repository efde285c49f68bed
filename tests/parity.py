"""Shared parity cases: the C ABI (through tiberate_fhe_b200.context.Tb200Context) against the oracle.

The same cases run
  * on a B200 through the CUDA library with torch tensors   (tests/test_gpu_parity.py, -m gpu), and
  * in the build container through tests/emu (the kernel sources compiled for the host) with NumPy
    arrays (tests/test_emu_parity.py) -- a check of the kernels' indexing logic, not a product path.
All comparisons are bit-exact (integer work).
"""

from __future__ import annotations

import numpy as np

from oracle import _c, call
from oracle.context import OracleContext, toy_primes
from oracle.engine import OracleEngine
from tiberate_fhe_b200._native import ExplicitConsts
from tiberate_fhe_b200.context import KeySwitchKeyView, Tb200Context, _ptr, galois_element


class Harness:
    """Moves arrays to the backend under test."""

    def __init__(self, lib, use_torch: bool, device: int = 0):
        self.lib, self.use_torch, self.device = lib, use_torch, device
        if use_torch:
            import torch

            self.torch = torch
            self.dev_name = f"cuda:{device}"

    def dev(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.int64)
        if self.use_torch:
            return self.torch.from_numpy(a).to(self.dev_name)
        return a.copy()

    def zeros(self, *shape):
        if self.use_torch:
            return self.torch.zeros(*shape, dtype=self.torch.int64, device=self.dev_name)
        return np.zeros(shape, dtype=np.int64)

    def host(self, t) -> np.ndarray:
        if self.use_torch:
            return t.detach().cpu().numpy()
        return np.asarray(t)

    def key(self, ksk, N):
        parts = [None if p is None else (self.dev(p[0]), self.dev(p[1])) for p in ksk]
        return KeySwitchKeyView(parts, N)


class Setup:
    def __init__(self, h: Harness, logN: int, q, K: int, seed: int = 0, rot_deltas=(1,)):
        self.h = h
        self.octx = OracleContext(logN, q, K)
        self.eng = OracleEngine(self.octx)
        self.ctx = Tb200Context(logN, q, K, device=h.device, lib=h.lib)
        self.rng = np.random.default_rng(seed)
        self.N = self.octx.N
        self.sk, _ = self.eng.gen_secret(self.rng)
        self.evk = self.eng.gen_evk(self.rng, self.sk)
        self.rotk = {d: self.eng.gen_rotk(self.rng, self.sk, d) for d in rot_deltas}
        self.evk_d = h.key(self.evk, self.N)
        self.rotk_d = {d: h.key(k, self.N) for d, k in self.rotk.items()}

    @classmethod
    def toy(cls, h, logN, num_scales, K, seed=0, scale_bits=40, **kw):
        return cls(h, logN, toy_primes(logN, num_scales, K, scale_bits=scale_bits), K, seed, **kw)

    def ct(self, level, batch=None):
        lp = self.octx.level_primes(level, False)
        if batch is None:
            return [self.eng.uniform(self.rng, lp), self.eng.uniform(self.rng, lp)]
        return [np.stack([self.eng.uniform(self.rng, lp) for _ in range(batch)]) for _ in range(2)]

    def close(self):
        self.ctx.close()


def far_primes(logN: int, num_scales: int, K: int) -> list[int]:
    """A toy chain whose 55-bit scale primes and 60-bit base / special primes lie FAR from a power of two
    (0.7 * 2^55, 0.75 * 2^60), unlike every prime of the reference's presets and of toy_primes: bounds of the
    lazy butterflies, offsets and reductions that only hold near 2^b would show here."""
    from oracle.context import is_prime

    def below(bound, count, skip=()):
        out, c = [], bound // (2 << logN)
        while len(out) < count:
            q = c * (2 << logN) + 1
            if q not in skip and is_prime(q):
                out.append(q)
            c -= 1
        return out

    return below(7 * (1 << 55) // 10, num_scales) + below(3 << 58, 1 + K)


def eq(h: Harness, got, want, what: str):
    g = h.host(got)
    w = np.asarray(want)
    assert g.shape == w.shape, f"{what}: shape {g.shape} != {w.shape}"
    if not np.array_equal(g, w):
        bad = np.argwhere(g != w)
        first = tuple(bad[0])
        raise AssertionError(
            f"{what}: {len(bad)} of {g.size} residues differ; first at {first}: got {g[first]} want {w[first]}"
        )


# ------------------------------------------------------------------------------------------------
def check_context(s: Setup):
    """Constants derived inside libtb200 == the oracle's restatement of the reference context."""
    pc = s.ctx.prime_consts()
    o = s.octx
    assert np.array_equal(pc[:, 0], o.qa) and np.array_equal(pc[:, 1], 2 * o.qa)
    assert np.array_equal(pc[:, 3], o.ka), "k = -q^-1 mod 2^62"
    assert np.array_equal(pc[:, 4], o.Rsa) and np.array_equal(pc[:, 5], _c(o.Rs_scale))
    assert np.array_equal(pc[:, 6], o.Ninva)
    for g in range(o.P):
        assert np.array_equal(s.ctx.twiddles(False, g), o.psi[g]), f"psi table prime {g}"
        assert np.array_equal(s.ctx.twiddles(True, g), o.ipsi[g]), f"ipsi table prime {g}"


def check_pointwise(s: Setup, level=0, with_special=False, signed_inputs=True):
    h, o, eng, ctx = s.h, s.octx, s.eng, s.ctx
    pr = o.level_primes(level, with_special)
    C, N = len(pr), s.N
    q, k = o.rows(pr)
    idx = np.asarray(pr)
    rng = s.rng
    a = eng.uniform(rng, pr)
    b = eng.uniform(rng, pr)
    if signed_inputs:  # lazy / signed operands as the reference meets them
        a = a + rng.integers(-1, 2, size=a.shape) * q[:, None] // 2
        b = b + rng.integers(0, 2, size=b.shape) * q[:, None]
    prime0 = pr[0]
    da, db = h.dev(a), h.dev(b)
    out = h.zeros(C, N)

    def ref2(name, *extra):
        r = np.empty_like(a)
        call(name, r, _c(a), _c(b), C, N, *extra)
        return r

    ctx.pointwise(0, da, db, out, prime0)
    eq(h, out, ref2("orc_mont_mult", q, k), "mont_mult")
    ctx.pointwise(1, da, db, out, prime0)
    eq(h, out, ref2("orc_mont_add", q), "mont_add")
    ctx.pointwise(2, da, db, out, prime0)
    eq(h, out, ref2("orc_mont_sub", q), "mont_sub")
    ctx.pointwise(3, da, db, out, prime0)
    eq(h, out, ref2("orc_mont_add_reduce_2q", q), "mont_add_reduce_2q")
    ctx.pointwise(4, da, db, out, prime0)
    eq(h, out, ref2("orc_mont_sub_reduce_2q", q), "mont_sub_reduce_2q")

    scal = _c([int(rng.integers(0, qq)) for qq in q.tolist()])
    dscal = h.dev(scal)

    def ref1(name, *extra):
        r = _c(a).copy()
        call(name, r, *extra)
        return r

    t = h.dev(a)
    ctx.pointwise(5, t, None, None, prime0, scal=dscal)
    eq(h, t, ref1("orc_mont_enter_scalar", scal, C, N, q, k), "mont_enter_scalar (in place)")
    t = h.dev(a)
    ctx.pointwise(6, t, None, None, prime0)
    eq(h, t, ref1("orc_mont_enter_scalar", o.Rsa[idx].copy(), C, N, q, k), "mont_enter_Rs")
    t = h.dev(a)
    ctx.pointwise(7, t, None, None, prime0)
    eq(h, t, ref1("orc_mont_enter_scalar", _c(o.Rs_scale)[idx].copy(), C, N, q, k), "mont_enter_Rs_scale")
    t = h.dev(a)
    ctx.pointwise(8, t, None, None, prime0)
    eq(h, t, ref1("orc_mont_reduce", C, N, q, k), "mont_reduce")
    t = h.dev(a)
    ctx.pointwise(9, t, None, None, prime0)
    eq(h, t, ref1("orc_reduce_2q", C, N, q), "reduce_2q")
    t = h.dev(a)
    ctx.pointwise(10, t, None, None, prime0)
    eq(h, t, ref1("orc_make_signed", C, N, q), "make_signed")
    t = h.dev(a)
    ctx.pointwise(11, t, None, None, prime0)
    eq(h, t, ref1("orc_make_unsigned", C, N, q), "make_unsigned")
    # tile_unsigned: [N] -> [C, N]
    small = rng.integers(-40, 41, size=N).astype(np.int64)
    r = np.empty((C, N), dtype=np.int64)
    call("orc_tile_unsigned", r, small, C, N, q)
    ctx.pointwise(12, h.dev(small), None, out, prime0)
    eq(h, out, r, "tile_unsigned")
    # pc_add_fused
    r = np.empty_like(a)
    call("orc_pc_add_fused", r, _c(a), _c(b), C, N, q, k, o.Rsa[idx].copy())
    ctx.pointwise(13, da, db, out, prime0)
    eq(h, out, r, "pc_add_fused")
    # legacy forms with explicit constant tensors (mont_enter, mont_add_legacy, tile_unsigned(_2q))
    half = (1 << 31) - 1
    ql, qh = h.dev(q & half), h.dev(q >> 31)
    kl, kh = h.dev(k & half), h.dev(k >> 31)
    two_q = h.dev(2 * q)
    ec = ExplicitConsts(_ptr(ql), _ptr(qh), _ptr(kl), _ptr(kh), 0)
    t = h.dev(a)
    ctx.pointwise(5, t, None, None, 0, scal=dscal, ec=ec)
    eq(h, t, ref1("orc_mont_enter_scalar", scal, C, N, q, k), "mont_enter (legacy, explicit constants)")
    ec2 = ExplicitConsts(0, 0, 0, 0, _ptr(two_q))
    ctx.pointwise(1, da, db, out, 0, ec=ec2)
    eq(h, out, ref2("orc_mont_add", q), "mont_add_legacy")
    # add_many
    K = 5
    st = np.stack([eng.uniform(rng, pr) + rng.integers(0, 2, size=(C, N)) * q[:, None] for _ in range(K)])
    for pairwise, name in ((False, "orc_mont_reduce_add_many_3d"), (True, "orc_mont_add_many_3d")):
        r = np.empty((C, N), dtype=np.int64)
        call(name, r, _c(st), K, C, N, q)
        ctx.add_many(h.dev(st), out, prime0, pairwise)
        eq(h, out, r, name)


def check_ntt(s: Setup, level=0, with_special=True, batch=2):
    h, o, eng, ctx = s.h, s.octx, s.eng, s.ctx
    pr = o.level_primes(level, with_special)
    N = s.N
    a = np.stack([eng.uniform(s.rng, pr) for _ in range(batch)])
    # a negative and a [q,2q) residue exercise the signed enter path
    a[0, 0, :8] -= o.q[pr[0]] // 3
    a[-1, -1, 8:16] += o.q[pr[-1]]
    want = np.stack([eng.enter_ntt(x, pr) for x in a])
    t = h.dev(a)
    ctx.ntt(t, pr[0], True)
    eq(h, t, want, "enter_ntt_radix2 (batched)")
    for mode, name in enumerate(("intt_radix2", "intt_radix2_exit", "intt_radix2_exit_reduce",
                                 "intt_radix2_exit_reduce_signed")):
        u = h.dev(want)
        ctx.intt(u, pr[0], mode)
        eq(h, u, np.stack([eng.intt(x, pr, mode) for x in want]), name)
    t2 = h.dev(a[0])
    ctx.ntt(t2, pr[0], False)
    eq(h, t2, eng.ntt(a[0], pr), "ntt_radix2")
    # negative residues must travel through the butterflies exactly as in the reference
    neg = want[0] - (o.qa[np.asarray(pr)] // 2)[:, None]
    u = h.dev(neg)
    ctx.intt(u, pr[0], 2)
    eq(h, u, eng.intt(neg, pr, 2), "intt_radix2_exit_reduce on negative lazy input")
    u = h.dev(neg)
    ctx.ntt(u, pr[0], False)
    eq(h, u, eng.ntt(neg, pr), "ntt_radix2 on negative lazy input")
    # sparse inputs: zero / tiny products put the Montgomery representative on the boundary that the
    # FP64 forward butterflies resolve through their integer fallback (ExactF64Pol)
    sp = np.zeros_like(a[0])
    sp[:, 1] = 1
    sp[:, 7] = o.qa[np.asarray(pr)] - 1
    sp[:, N // 2] = 2
    for enter in (True, False):
        u = h.dev(sp)
        ctx.ntt(u, pr[0], enter)
        eq(h, u, eng.enter_ntt(sp, pr) if enter else eng.ntt(sp, pr), f"forward NTT of a sparse polynomial (enter={enter})")
    u = h.dev(np.zeros_like(a[0]))
    ctx.ntt(u, pr[0], True)
    eq(h, u, eng.enter_ntt(np.zeros_like(a[0]), pr), "forward NTT of zero")
    # beyond the FP64 domain (|x| >= 2^50): the tile falls back to the integer butterflies
    big = a[0].copy()
    big[:, 3] += np.int64(1) << 52
    u = h.dev(big)
    ctx.ntt(u, pr[0], False)
    eq(h, u, eng.ntt(big, pr), "ntt_radix2 on an input beyond 2^50")
    # unreduced sums of lazy values on the 40-bit limbs (documented domain of the mod-q route: |x| < 2^51)
    wide = want[0].copy()
    for r, g in enumerate(pr):
        if o.q[g] < (1 << 42):
            wide[r, 0:N:7] += 300 * o.q[g]
            wide[r, 3:N:11] -= 411 * o.q[g]
    u = h.dev(wide)
    ctx.intt(u, pr[0], 2)
    eq(h, u, eng.intt(wide, pr, 2), "intt_radix2_exit_reduce on unreduced lazy input")
    # products that land in the band where the Montgomery representative depends on floor(O S / 2^62): the first
    # stage pairs j with j + N/2 under one twiddle S = w R, so O = r w^-1 mod q gives MM(S, O) = r (mod q) with
    # r tiny -- half of them as the lazy value O + q (the deferred-reduction forward path, ExactSumPol, must
    # follow the reference bit for bit through its integer fallback)
    band = np.stack([s.rng.integers(0, 2 * int(o.q[g]), size=N, dtype=np.int64) for g in pr])
    Rinv = [pow(1 << 62, -1, int(o.q[g])) for g in pr]
    for r, g in enumerate(pr):
        qg = int(o.q[g])
        w = int(ctx.twiddles(False, g)[1]) * Rinv[r] % qg  # plain twiddle of stage 0
        winv = pow(w, -1, qg)
        small = s.rng.integers(0, 1 << 21, size=N // 2)
        vals = [(int(v) * winv) % qg + (qg if (i & 1) else 0) for i, v in enumerate(small)]
        band[r, N // 2:] = np.array(vals, dtype=np.int64)
    u = h.dev(band)
    ctx.ntt(u, pr[0], False)
    eq(h, u, eng.ntt(band, pr), "ntt_radix2 with first-stage products on the representative boundary")
    # row-offset view (the reference's rescale outputs are views with storage offset N)
    big = h.dev(np.concatenate([a[0][:1], a[0]], axis=0))
    view = big[1:]
    ctx.ntt(view, pr[0], True)
    eq(h, view, want[0], "enter_ntt_radix2 on a row-offset view")
    # `rows` smaller than the tensor: ntt_radix2 skips the trailing special rows (A.0 exception)
    if with_special and o.K < len(pr):
        u = h.dev(a[0])
        ctx.ntt(u, pr[0], False, rows=len(pr) - o.K)
        w = a[0].copy()
        w[: len(pr) - o.K] = eng.ntt(a[0][: len(pr) - o.K], pr[: len(pr) - o.K])
        eq(h, u, w, "ntt_radix2 on the ordinary rows only")


def check_he_ops(s: Setup, level=0):
    """Op-layer fused kernels: rescale_rows, extend, codec_rotate, divide_by_p."""
    h, o, eng, ctx = s.h, s.octx, s.eng, s.ctx
    N = s.N
    rng = s.rng
    lp = o.level_primes(level, False)
    L = len(lp)
    if L > 1:
        a = eng.uniform(rng, lp)
        t = h.dev(a)
        scales = o.rescale_scales[level][: L - 1].copy()
        for exact in (True, False):
            t = h.dev(a)
            ctx.rescale_rows(t[1:], level + 1, h.dev(scales), t[0], o.q[level] // 2, exact)
            eq(h, t[1:], eng.rescale_poly(a, level, exact), f"rescale_rows exact={exact}")
    # extend of every group at this level
    a = eng.uniform(rng, lp)
    tgt = o.level_primes(level, True)
    for g, primes in o.part.level_groups(level):
        state = eng.pre_extend(a, level, primes)
        _, _, Lenter = o.group_scalars(primes)
        out = h.zeros(len(tgt), N)
        le = h.dev(Lenter) if Lenter.shape[0] else None
        ctx.extend(len(tgt), level, h.dev(state), le, level, out)
        eq(h, out, eng.extend(state, level, primes), f"extend group {g}")
    # codec_rotate
    for delta in (1, 5, N // 2 - 1):
        perm = eng.galois_perm(delta)
        q, _ = o.rows(lp)
        out = h.zeros(L, N)
        ctx.codec_rotate(h.dev(a), h.dev(perm), h.dev(2 * q), out)
        eq(h, out, eng.codec_rotate(a, level, perm), f"codec_rotate delta={delta}")
    # divide_by_p
    d = eng.uniform(rng, tgt)
    td = h.dev(d)
    out = h.zeros(L, N)
    ctx.divide_by_p(level, td[:L], td[L:], out)
    eq(h, out, eng.divide_by_p(d, level), "create_switcher_divide_by_p")


# Internal-transform modes of the fused engine calls, all bit-identical by contract:
# (mod-q path?, eighths of the 40-bit limbs on the FP64 pipe)
# optional third entry: 0 = the key-switch core as three kernels instead of the fused one
# and fourth: 1 = ModDown inside the inverse pass-A exit instead of separate kernels
ENGINE_MODES = ((False, 0), (True, 0), (True, 3), (True, 8, 0, 1), (True, 8))  # the default last


def set_mode(s: "Setup", mode):
    fast, share = mode[0], mode[1]
    s.fast_mode = bool(fast)
    s.ctx.set_fast(fast)
    s.ctx.set_f64_share(share)
    s.ctx.set_tuning(s.ctx.TUNE_FUSED_CORE, mode[2] if len(mode) > 2 else 1)
    s.ctx.set_tuning(s.ctx.TUNE_FUSED_MODDOWN, mode[3] if len(mode) > 3 else 0)


def check_engine(s: Setup, level: int, batch=None, ops=("rescale", "keyswitch", "switch_key", "rotate", "hoisted", "cc_mult",
                                                        "triplet", "pc_mult", "addsub")):
    h, o, eng, ctx = s.h, s.octx, s.eng, s.ctx
    N = s.N
    L = o.num_ordinary - level
    B = batch or 1
    ct1, ct2 = s.ct(level, batch), s.ct(level, batch)
    d1 = [h.dev(x) for x in ct1]
    d2 = [h.dev(x) for x in ct2]

    def each(fn):
        """Apply an oracle function per batch entry and stack."""
        if batch is None:
            return fn(ct1, ct2)
        outs = [fn([ct1[0][b], ct1[1][b]], [ct2[0][b], ct2[1][b]]) for b in range(B)]
        return [np.stack([o_[i] for o_ in outs]) for i in range(len(outs[0]))]

    shp = (L, N) if batch is None else (B, L, N)
    shp1 = (L - 1, N) if batch is None else (B, L - 1, N)
    can_rescale = L > 1
    if "rescale" in ops and can_rescale:
        o0, o1 = h.zeros(*shp1), h.zeros(*shp1)
        ctx.rescale(level, d1[0], d1[1], o0, o1)
        w = each(lambda a, b: eng.rescale(a, level))
        eq(h, o0, w[0], f"rescale level {level} c0")
        eq(h, o1, w[1], f"rescale level {level} c1")
    if "keyswitch" in ops:
        o0, o1 = h.zeros(*shp), h.zeros(*shp)
        ctx.keyswitch(level, d1[1], s.evk_d, o0, o1)
        w = each(lambda a, b: list(eng.create_switcher(a[1], s.evk, level)))
        eq(h, o0, w[0], f"create_switcher level {level} out0")
        eq(h, o1, w[1], f"create_switcher level {level} out1")
    if "switch_key" in ops:
        o0, o1 = h.zeros(*shp), h.zeros(*shp)
        ctx.switch_key(level, d1[0], d1[1], s.evk_d, o0, o1)
        w = each(lambda a, b: eng.switch_key(a, s.evk, level))
        eq(h, o0, w[0], f"switch_key level {level} c0")
        eq(h, o1, w[1], f"switch_key level {level} c1")
    if "rotate" in ops:
        for delta, rk in s.rotk.items():
            o0, o1 = h.zeros(*shp), h.zeros(*shp)
            ctx.rotate(level, galois_element(N, delta), d1[0], d1[1], s.rotk_d[delta], o0, o1)
            w = each(lambda a, b: eng.rotate_single(a, rk, delta, level))
            eq(h, o0, w[0], f"rotate_single delta {delta} level {level} c0")
            eq(h, o1, w[1], f"rotate_single delta {delta} level {level} c1")
            ctx.rotate(level, galois_element(N, delta), d1[0], d1[1], None, o0, o1)
            perm = eng.galois_perm(delta)
            w = each(lambda a, b: [eng.codec_rotate(a[0], level, perm), eng.codec_rotate(a[1], level, perm)])
            eq(h, o0, w[0], f"automorphism only delta {delta} c0")
    if "hoisted" in ops and s.rotk and getattr(s, "fast_mode", True):  # hoisting exists on the mod-q path only
        deltas = list(s.rotk)
        R = len(deltas)
        o0, o1 = h.zeros(R, *shp), h.zeros(R, *shp)
        ctx.rotate_hoisted(level, [galois_element(N, d) for d in deltas], d1[0], d1[1], [s.rotk_d[d] for d in deltas], o0, o1)
        g0, g1 = h.host(o0), h.host(o1)
        for b in range(B):
            one = [ct1[0], ct1[1]] if batch is None else [ct1[0][b], ct1[1][b]]
            want = eng.rotate_hoisted(one, s.rotk, deltas, level)
            for r, d in enumerate(deltas):
                got = (g0[r], g1[r]) if batch is None else (g0[r][b], g1[r][b])
                assert np.array_equal(got[0], want[r][0]), f"rotate_hoisted delta {d} level {level} c0 (batch entry {b})"
                assert np.array_equal(got[1], want[r][1]), f"rotate_hoisted delta {d} level {level} c1 (batch entry {b})"
    if "cc_mult" in ops:
        for pre in ((True, False) if can_rescale else (False,)):
            sh = shp1 if pre else shp
            o0, o1 = h.zeros(*sh), h.zeros(*sh)
            ctx.cc_mult_relin(level, d1[0], d1[1], d2[0], d2[1], s.evk_d, o0, o1, pre)
            w = each(lambda a, b: eng.cc_mult(a, b, s.evk, level, pre_rescale=pre)[0])
            eq(h, o0, w[0], f"cc_mult+relin level {level} pre_rescale={pre} c0")
            eq(h, o1, w[1], f"cc_mult+relin level {level} pre_rescale={pre} c1")
        # squaring: the same tensors as both operands (two transforms instead of four)
        pre = can_rescale
        sh = shp1 if pre else shp
        o0, o1 = h.zeros(*sh), h.zeros(*sh)
        ctx.cc_mult_relin(level, d1[0], d1[1], d1[0], d1[1], s.evk_d, o0, o1, pre)
        w = each(lambda a, b: eng.cc_mult(a, a, s.evk, level, pre_rescale=pre)[0])
        eq(h, o0, w[0], f"cc_mult+relin of a ciphertext with itself, level {level} c0")
        eq(h, o1, w[1], f"cc_mult+relin of a ciphertext with itself, level {level} c1")
    if "triplet" in ops:
        pre = can_rescale
        sh = shp1 if pre else shp
        lvl = level + (1 if pre else 0)
        t = [h.zeros(*sh) for _ in range(3)]
        ctx.cc_mult_triplet(level, d1[0], d1[1], d2[0], d2[1], t[0], t[1], t[2], pre)
        w = each(lambda a, b: eng.cc_mult(a, b, s.evk, level, pre_rescale=pre, post_relin=False)[0])
        for i in range(3):
            eq(h, t[i], w[i], f"cc_mult triplet d{i} level {level}")
        o0, o1 = h.zeros(*sh), h.zeros(*sh)
        ctx.relinearize(lvl, t[0], t[1], t[2], s.evk_d, o0, o1)
        if batch is None:
            wr = eng.relinearize(w, s.evk, lvl)
        else:
            rs = [eng.relinearize([w[0][b], w[1][b], w[2][b]], s.evk, lvl) for b in range(B)]
            wr = [np.stack([r[i] for r in rs]) for i in range(2)]
        eq(h, o0, wr[0], f"relinearize level {lvl} c0")
        eq(h, o1, wr[1], f"relinearize level {lvl} c1")
        for i in range(3):  # inputs untouched
            eq(h, t[i], w[i], f"relinearize must not modify its triplet (d{i})")
    if "pc_mult" in ops:
        lp = o.level_primes(level, False)
        pt = eng.enter_ntt(eng.uniform(s.rng, lp), lp)
        for post in ((True, False) if can_rescale else (False,)):
            sh = shp1 if post else shp
            o0, o1 = h.zeros(*sh), h.zeros(*sh)
            ctx.pc_mult(level, h.dev(pt), d1[0], d1[1], o0, o1, post)
            w = each(lambda a, b: eng.pc_mult(pt, a, level, post_rescale=post)[0])
            eq(h, o0, w[0], f"pc_mult level {level} post_rescale={post} c0")
            eq(h, o1, w[1], f"pc_mult level {level} post_rescale={post} c1")
    if "addsub" in ops:
        o0, o1 = h.zeros(*shp), h.zeros(*shp)
        ctx.cc_addsub(level, False, d1[0], d1[1], d2[0], d2[1], o0, o1)
        w = each(lambda a, b: eng.cc_add(a, b, level))
        eq(h, o0, w[0], "cc_add c0")
        eq(h, o1, w[1], "cc_add c1")
        ctx.cc_addsub(level, True, d1[0], d1[1], d2[0], d2[1], o0, o1)
        w = each(lambda a, b: eng.cc_sub(a, b, level))
        eq(h, o0, w[0], "cc_sub c0")
        eq(h, o1, w[1], "cc_sub c1")
    # inputs must be left untouched by every engine call
    eq(h, d1[0], ct1[0], "engine calls must not modify their inputs")
    eq(h, d1[1], ct1[1], "engine calls must not modify their inputs")
