"""Oracle restatement of the hot CkksEngine methods on NumPy int64 arrays (single device).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each method cites the reference lines it
follows step by step; the element-wise work is done by oracle/ckks_oracle.c.

Data layout is the reference's: a polynomial is int64 [limbs, N], row i =
residues modulo prime level+i; a ciphertext is (c0, c1); at rest: coefficient domain, canonical.
A key-switch key is a list over *global digit-group id* of (b, a), each [P, N] NTT+Montgomery.
"""

from __future__ import annotations

import numpy as np

from . import _c, call
from .context import OracleContext


class OracleEngine:
    def __init__(self, ctx: OracleContext):
        self.ctx = ctx

    # ------------------------------------------------------------------ primitive wrappers
    def _qk(self, level, with_special, rows=None):
        pr = self.ctx.level_primes(level, with_special)
        if rows is not None:
            assert len(pr) == rows, (len(pr), rows)
        return pr, *self.ctx.rows(pr)

    def mont_mult(self, a, b, primes):
        q, k = self.ctx.rows(primes)
        out = np.empty_like(a)
        call("orc_mont_mult", out, _c(a), _c(b), a.shape[0], a.shape[1], q, k)
        return out

    def enter_ntt(self, a, primes):
        """ntt_radix2_cuda.cu:98-136: MM(a, Rs) then the logN stages. In place on a copy."""
        a = _c(a).copy()
        q, k = self.ctx.rows(primes)
        idx = np.asarray(primes)
        call("orc_mont_enter_scalar", a, self.ctx.Rsa[idx].copy(), a.shape[0], a.shape[1], q, k)
        call("orc_ntt_stages", a, a.shape[0], a.shape[1], _c(self.ctx.psi[idx]), q, k)
        return a

    def ntt(self, a, primes):
        a = _c(a).copy()
        q, k = self.ctx.rows(primes)
        call("orc_ntt_stages", a, a.shape[0], a.shape[1], _c(self.ctx.psi[np.asarray(primes)]), q, k)
        return a

    def intt(self, a, primes, mode):
        """mode 0 intt_radix2, 1 _exit, 2 _exit_reduce, 3 _exit_reduce_signed."""
        a = _c(a).copy()
        q, k = self.ctx.rows(primes)
        idx = np.asarray(primes)
        call("orc_intt_stages", a, a.shape[0], a.shape[1], _c(self.ctx.ipsi[idx]), q, k)
        call("orc_intt_epilogue", a, a.shape[0], a.shape[1], self.ctx.Ninva[idx].copy(), q, k, mode)
        return a

    # ------------------------------------------------------------------ rescale
    def rescale_poly(self, a, level, exact=True):
        """ckks_engine.py:1520-1618 for one polynomial [L+1, N] at `level` -> [L, N]."""
        ctx = self.ctx
        a = _c(a)
        keep = a[1:].copy()
        primes = list(range(level + 1, level + a.shape[0]))
        q, k = ctx.rows(primes)
        scales = ctx.rescale_scales[level][: len(primes)].copy()
        call("orc_rescale", keep, scales, a[0].copy(), ctx.q[level] // 2, 1 if exact else 0,
             keep.shape[0], keep.shape[1], q, k)
        return keep

    def rescale(self, ct, level, exact=True):
        return [self.rescale_poly(ct[0], level, exact), self.rescale_poly(ct[1], level, exact)]

    # ------------------------------------------------------------------ key switch
    def pre_extend(self, a, level, primes):
        """ckks_engine.py:863-924: mixed-radix digits of one digit group.
        `a` is [L, N] at `level`; `primes` are the group's global prime ids."""
        ctx = self.ctx
        alpha = len(primes)
        rows = [p - level for p in primes]
        a_part = _c(a[rows[0]: rows[-1] + 1])
        state = np.repeat(a_part[0:1], alpha, axis=0).copy()
        Y_scalar, L_scalar, _ = ctx.group_scalars(primes)
        for i in range(alpha - 1):
            q1, k1 = ctx.rows([primes[i + 1]])
            Y = (a_part[i + 1] - state[i + 1])[None, :].copy()
            call("orc_mont_enter_scalar", Y, _c([Y_scalar[i]]), 1, Y.shape[1], q1, k1)
            state[i + 1] = Y[0]
            if i + 2 < alpha:
                qs, ks = ctx.rows(primes[i + 2:])
                new_state = np.repeat(Y, alpha - (i + 2), axis=0).copy()
                call("orc_mont_enter_scalar", new_state, _c(L_scalar[i]), new_state.shape[0],
                     new_state.shape[1], qs, ks)
                state[i + 2:] += new_state
        return state

    def extend(self, state, level, primes):
        """ckks_engine.py:976-1012 + he_fused_cuda.cu:276-312: digits -> all L+K limbs (Montgomery)."""
        ctx = self.ctx
        tgt = ctx.level_primes(level, True)
        q, k = ctx.rows(tgt)
        _, _, Lenter = ctx.group_scalars(primes)
        le = _c(Lenter[:, level:]) if Lenter.shape[0] else np.zeros((0, len(tgt)), dtype=np.int64)
        out = np.empty((len(tgt), state.shape[1]), dtype=np.int64)
        call("orc_extend", out, _c(state), state.shape[0], le, len(tgt), state.shape[1],
             ctx.Rsa[np.asarray(tgt)].copy(), q, k)
        return out

    def create_switcher(self, a, ksk, level):
        """ckks_engine.py:1201-1363.  a: [L, N] coefficient/canonical at `level`;
        returns (out0, out1), each [L, N] canonical."""
        ctx = self.ctx
        tgt = ctx.level_primes(level, True)
        parts0, parts1 = [], []
        order = {g: pr for g, pr in ctx.part.level_groups(level)}
        for g in ctx.part.storage_order(level):
            primes = order[g]
            state = self.pre_extend(a, level, primes)
            ext = self.extend(state, level, primes)
            ext = self.ntt(ext, tgt)  # :1381
            b, aa = ksk[g]
            parts0.append(self.mont_mult(ext, _c(b[level:]), tgt))  # :1392-1397
            parts1.append(self.mont_mult(ext, _c(aa[level:]), tgt))
        q, _ = ctx.rows(tgt)
        d = []
        for parts in (parts0, parts1):
            st = _c(np.stack(parts))
            acc = np.empty_like(st[0])
            call("orc_mont_reduce_add_many_3d", acc, st, st.shape[0], st.shape[1], st.shape[2], q)  # :1323
            d.append(self.intt(acc, tgt, 2))  # :1327-1328
        return self.divide_by_p(d[0], level), self.divide_by_p(d[1], level)

    def divide_by_p(self, d, level):
        """ckks_engine.py:1330-1360 + he_fused_cuda.cu:433-584 (ModDown)."""
        ctx = self.ctx
        K = ctx.K
        c = _c(d[:-K])
        p = _c(d[-K:]).copy()
        ordp = ctx.level_primes(level, False)
        q, k = ctx.rows(ordp)
        qsp, ksp = ctx.rows(ctx.sp)
        pir_sp = _c(ctx.PiR[:, ctx.num_ordinary:])  # [k][row]
        call("orc_chain_backward", p, K, p.shape[1], pir_sp, qsp, ksp)
        pir = _c(ctx.PiR[:, level: ctx.num_ordinary])
        out = np.empty_like(c)
        call("orc_divide_by_p", out, c, p, K, pir, c.shape[0], c.shape[1],
             ctx.Rsa[np.asarray(ordp)].copy(), q, k)
        return out

    def switch_key(self, ct, ksk, level):
        """ckks_engine.py:1403-1420."""
        d0, d1 = self.create_switcher(ct[1], ksk, level)
        q, _ = self.ctx.rows(self.ctx.level_primes(level, False))
        new0 = np.empty_like(d0)
        call("orc_mont_add_reduce_2q", new0, _c(ct[0]), d0, d0.shape[0], d0.shape[1], q)
        return [new0, d1]

    # ------------------------------------------------------------------ multiplication
    def cc_mult(self, a, b, evk, level, pre_rescale=True, post_relin=True):
        """ckks_engine.py:1640-1692.  `level` is the level of the inputs."""
        if pre_rescale:
            x, y = self.rescale(a, level), self.rescale(b, level)
            level += 1
        else:
            x, y = a, b
        pr = self.ctx.level_primes(level, False)
        x0, x1, y0, y1 = (self.enter_ntt(t, pr) for t in (x[0], x[1], y[0], y[1]))
        d0 = self.mont_mult(x0, y0, pr)
        x0y1 = self.mont_mult(x0, y1, pr)
        x1y0 = self.mont_mult(x1, y0, pr)
        q, _ = self.ctx.rows(pr)
        d1 = np.empty_like(d0)
        call("orc_mont_add", d1, x0y1, x1y0, d0.shape[0], d0.shape[1], q)
        d2 = self.mont_mult(x1, y1, pr)
        if not post_relin:
            return [d0, d1, d2], level
        return self.relinearize([d0, d1, d2], evk, level), level

    def relinearize(self, triplet, evk, level):
        """ckks_engine.py:1695-1732."""
        pr = self.ctx.level_primes(level, False)
        d0, d1, d2 = (self.intt(t, pr, 2) for t in triplet)
        s0, s1 = self.create_switcher(d2, evk, level)
        d0 = d0 + s0
        d1 = d1 + s1
        q, _ = self.ctx.rows(pr)
        call("orc_reduce_2q", d0, d0.shape[0], d0.shape[1], q)
        call("orc_reduce_2q", d1, d1.shape[0], d1.shape[1], q)
        return [d0, d1]

    # ------------------------------------------------------------------ rotation
    def galois_perm(self, delta):
        """ckks_engine.py:1817-1821 + utils/encoding.py:71-88."""
        N = self.ctx.N
        d = delta % N
        leap = (3 ** d - 1) // 2 % (2 * N)
        p = 2 * leap + 1
        return (p * np.arange(N, dtype=np.int64)) % (2 * N)

    def codec_rotate(self, a, level, perm):
        pr = self.ctx.level_primes(level, False)[: a.shape[0]]
        q, _ = self.ctx.rows(pr)
        out = np.empty_like(a)
        call("orc_codec_rotate", out, _c(a), _c(perm), a.shape[0], a.shape[1], q)
        return out

    def rotate_single(self, ct, rotk, delta, level):
        """ckks_engine.py:1804-1840."""
        perm = self.galois_perm(delta)
        rot = [self.codec_rotate(_c(ct[0]), level, perm), self.codec_rotate(_c(ct[1]), level, perm)]
        return self.switch_key(rot, rotk, level)

    def ntt_slot_perm(self, galois):
        """X -> X^g in the NTT domain: slot i holds the evaluation at psi^(2 brev(i) + 1), and
        (sigma_g a)(psi^e) = a(psi^(e g)), so slot i of sigma_g(a) is slot brev(((2 brev(i)+1) g mod 2N - 1)/2) of a."""
        from .context import bit_reverse

        N, logN = self.ctx.N, self.ctx.logN
        br = np.array([bit_reverse(i, logN) for i in range(N)], dtype=np.int64)
        e = ((2 * br + 1) * int(galois)) % (2 * N)
        return br[(e - 1) // 2]  # brev is an involution: br[x] = brev(x)

    def rotate_hoisted(self, ct, rotks, deltas, level):
        """Hoisted rotations (an extension beyond the reference; restates tb200_rotate_hoisted): the digits,
        extension and forward transform of c1 are computed ONCE (as in create_switcher, ckks_engine.py:1201-1400),
        each rotation applies the automorphism to the transformed extension (a slot permutation), multiplies by its
        own key, and finishes as switch_key does (:1316-1363, :1403-1420) with the rotated c0."""
        ctx = self.ctx
        tgt = ctx.level_primes(level, True)
        order = {g: pr for g, pr in ctx.part.level_groups(level)}
        exts = []
        for g in ctx.part.storage_order(level):
            state = self.pre_extend(_c(ct[1]), level, order[g])
            exts.append((g, self.ntt(self.extend(state, level, order[g]), tgt)))
        q, _ = ctx.rows(tgt)
        qo, _ = ctx.rows(ctx.level_primes(level, False))
        outs = []
        for delta in deltas:
            d = delta % ctx.N
            gal = (2 * ((3 ** d - 1) // 2 % (2 * ctx.N)) + 1) % (2 * ctx.N)
            pi = self.ntt_slot_perm(gal)
            ksk = rotks[delta]
            halves = []
            for which in (0, 1):
                parts = [self.mont_mult(_c(ext[:, pi]), _c(ksk[g][which][level:]), tgt) for g, ext in exts]
                st = _c(np.stack(parts))
                acc = np.empty_like(st[0])
                call("orc_mont_reduce_add_many_3d", acc, st, st.shape[0], st.shape[1], st.shape[2], q)
                halves.append(self.divide_by_p(self.intt(acc, tgt, 2), level))
            r0 = self.codec_rotate(_c(ct[0]), level, self.galois_perm(delta))
            new0 = np.empty_like(r0)
            call("orc_mont_add_reduce_2q", new0, r0, halves[0], r0.shape[0], r0.shape[1], qo)
            outs.append([new0, halves[1]])
        return outs

    # ------------------------------------------------------------------ add / plaintext mult
    def cc_add(self, a, b, level):
        """ckks_engine.py:1932-1956."""
        q, _ = self.ctx.rows(self.ctx.level_primes(level, False))
        out = []
        for x, y in zip(a, b):
            o = np.empty_like(x)
            call("orc_mont_add_reduce_2q", o, _c(x), _c(y), x.shape[0], x.shape[1], q)
            out.append(o)
        return out

    def cc_sub(self, a, b, level):
        q, _ = self.ctx.rows(self.ctx.level_primes(level, False))
        out = []
        for x, y in zip(a, b):
            o = np.empty_like(x)
            call("orc_mont_sub_reduce_2q", o, _c(x), _c(y), x.shape[0], x.shape[1], q)
            out.append(o)
        return out

    def pc_mult(self, pt_ntt, ct, level, post_rescale=True):
        """ckks_engine.py:2542-2580; pt_ntt is the cached NTT+Montgomery plaintext [L, N]."""
        pr = self.ctx.level_primes(level, False)
        out = []
        for c in ct:
            x = self.enter_ntt(c, pr)
            x = self.mont_mult(_c(pt_ntt), x, pr)
            out.append(self.intt(x, pr, 2))
        if post_rescale:
            return self.rescale(out, level), level + 1
        return out, level

    # ------------------------------------------------------------------ key material for tests
    # The reference draws from its CSPRNG extension (out of scope); here the same
    # compositions are fed from a NumPy generator so that tests own valid keys.
    def gen_secret(self, rng):
        """ckks_engine.py:486-504: ternary, tile_unsigned over all P primes, enter_ntt."""
        ctx = self.ctx
        s = rng.integers(-1, 2, size=ctx.N, dtype=np.int64)
        allp = list(range(ctx.P))
        q, _ = ctx.rows(allp)
        t = np.empty((ctx.P, ctx.N), dtype=np.int64)
        call("orc_tile_unsigned", t, s, ctx.P, ctx.N, q)
        return self.enter_ntt(t, allp), s

    def gen_error(self, rng, sigma=3.2):
        return np.rint(rng.normal(0.0, sigma, size=self.ctx.N)).astype(np.int64)

    def uniform(self, rng, primes):
        return np.stack([rng.integers(0, self.ctx.q[p], size=self.ctx.N, dtype=np.int64) for p in primes])

    def gen_public(self, rng, sk, with_special=True):
        """ckks_engine.py:507-557: pk = (e - a*sk, a) over level 0."""
        ctx = self.ctx
        pr = ctx.level_primes(0, with_special)
        q, _ = ctx.rows(pr)
        e = np.empty((len(pr), ctx.N), dtype=np.int64)
        call("orc_tile_unsigned", e, self.gen_error(rng), len(pr), ctx.N, q)
        e = self.enter_ntt(e, pr)
        a = self.uniform(rng, pr)
        sa = self.mont_mult(a, _c(sk[: len(pr)]), pr)
        pk0 = np.empty_like(sa)
        call("orc_mont_sub", pk0, e, sa, sa.shape[0], sa.shape[1], q)
        return pk0, a

    def gen_ksk(self, rng, sk_from, sk_to):
        """ckks_engine.py:796-860: for every digit group G, pk_G.b[rows of G] += P*R*sk_from."""
        ctx = self.ctx
        no = ctx.num_ordinary
        Psk = _c(sk_from[:no]).copy()
        ordp = ctx.level_primes(0, False)
        q, k = ctx.rows(ordp)
        call("orc_mont_enter_scalar", Psk, ctx.mont_PR, no, ctx.N, q, k)
        ksk = [None] * (ctx.part.num_partitions + 1)
        for g, primes in ctx.part.level_groups(0):
            b, a = self.gen_public(rng, sk_to, True)
            lo, hi = primes[0], primes[-1] + 1
            qq, _ = ctx.rows(primes)
            upd = np.empty((hi - lo, ctx.N), dtype=np.int64)
            call("orc_mont_add", upd, _c(b[lo:hi]), _c(Psk[lo:hi]), hi - lo, ctx.N, qq)
            b[lo:hi] = upd
            ksk[g] = (b, a)
        return ksk

    def gen_evk(self, rng, sk):
        allp = list(range(self.ctx.P))
        return self.gen_ksk(rng, self.mont_mult(sk, sk, allp), sk)

    def rotate_plain(self, m, delta):
        """utils/encoding.py:275-291 on a [C, N] array (no modular fix-up)."""
        perm = self.galois_perm(delta)
        N = self.ctx.N
        out = np.zeros_like(m)
        sign = np.where((perm // N) % 2 == 1, -1, 1).astype(np.int64)
        out[:, perm % N] = m * sign[None, :]
        return out

    def gen_rotk(self, rng, sk, delta):
        """ckks_engine.py:1739-1764: sk -> coefficient (intt, stays Montgomery), permute, ntt."""
        ctx = self.ctx
        no = ctx.num_ordinary
        ordp = ctx.level_primes(0, False)
        s = self.intt(_c(sk[:no]), ordp, 0)
        s = self.rotate_plain(s, delta)
        s = self.ntt(s, ordp)
        return self.gen_ksk(rng, s, sk)

    def encrypt_poly(self, rng, m, pk, level=0):
        """Minimal encryptor for tests (message polynomial `m` [N] small signed ints already
        scaled): ct = (v*pk0 + m + e0, v*pk1 + e1), coefficient domain canonical.
        Follows ckks_engine.py:565-636 without the scale multiplication."""
        ctx = self.ctx
        pr = ctx.level_primes(level, False)
        q, _ = ctx.rows(pr)
        L = len(pr)

        def tile(x):
            t = np.empty((L, ctx.N), dtype=np.int64)
            call("orc_tile_unsigned", t, _c(x), L, ctx.N, q)
            return t

        v = self.enter_ntt(tile(rng.integers(0, 2, size=ctx.N, dtype=np.int64)), pr)
        out = []
        for pkx, add in ((pk[0], m + self.gen_error(rng)), (pk[1], self.gen_error(rng))):
            x = self.mont_mult(v, _c(pkx[level: level + L]), pr)
            x = self.intt(x, pr, 1)
            o = np.empty_like(x)
            call("orc_mont_add_reduce_2q", o, x, tile(add), L, ctx.N, q)
            out.append(o)
        return out

    def decrypt_poly(self, ct, sk, level):
        """c0 + c1*s as centred big ints (CRT over the live ordinary primes); tests only."""
        ctx = self.ctx
        pr = ctx.level_primes(level, False)
        q, _ = ctx.rows(pr)
        a = self.enter_ntt(ct[1], pr)
        sa = self.intt(self.mont_mult(a, _c(sk[level: level + len(pr)]), pr), pr, 1)
        pt = np.empty_like(sa)
        call("orc_mont_add_reduce_2q", pt, _c(ct[0]), sa, sa.shape[0], sa.shape[1], q)
        return crt_centered(pt, [ctx.q[p] for p in pr])


def crt_centered(res, primes):
    """[L, N] canonical residues -> list of N centred Python ints."""
    import math

    Q = math.prod(primes)
    out = [0] * res.shape[1]
    for r, qi in zip(res, primes):
        Qi = Q // qi
        c = Qi * pow(Qi, -1, qi)
        for j, v in enumerate(r.tolist()):
            out[j] += (v % qi) * c
    return [((x % Q) + Q // 2) % Q - Q // 2 for x in out]
