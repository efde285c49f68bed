"""Key generation, encryption and decryption sequencing: the reference's
tiberate/ckks_engine.py:486-557 (_create_secret_key, _create_public_key), :565-637 (encrypt),
:707-760 (decrypt_double), :640-705 (decrypt_triplet), :796-860 (create_key_switching_key),
:1621-1634 (_create_evk), :1739-1764 (_create_rotation_key), :241-291 (mont_PR, final scalars),
composed from the same operators in the same order with the same CSPRNG consumption, so that an
engine seeded like a reference engine produces bit-identical keys and ciphertexts
(tests/test_gpu_keygen.py runs both side by side).

Single device per engine: every per-device list has one entry.  None of this is on the hot path --
each step is one op-layer call (wrapper/*.py -> C ABI); the random draws are the CSPRNG kernels.
"""

from __future__ import annotations

import math

import torch

from .context import galois_element
from .typing import FLAGS, Ciphertext, EvaluationKey, KeySwitchKey, PublicKey, RotationKey, SecretKey
from . import wrapper


class _RotationKeys(dict):
    """engine.rotk[delta]: created on first use (the reference's CachedDict, ckks_engine.py:380-388)."""

    def __init__(self, make):
        super().__init__()
        self._make = make

    def __missing__(self, delta):
        self[delta] = k = self._make(delta)
        return k


class KeyGenMixin:
    # ---- NTTContext method equivalents (tiberate/context/ntt_context.py:715-871) ------------------
    def _sp(self, mult_type):
        return self.ctx.K if mult_type == -1 else 0

    def _primes(self, lvl, mult_type):
        """prime indices of a tensor at `lvl` (mult_type -1: ordinary rows, -2: with the special primes)"""
        end = self.ctx.P if mult_type == -2 else self.ctx.num_ordinary
        return list(range(lvl, end))

    def _two_q(self, lvl, mult_type):
        key = ("2q", lvl, mult_type)
        if key not in self._consts:
            self._consts[key] = torch.tensor([2 * self.ctx.q[i] for i in self._primes(lvl, mult_type)],
                                             dtype=torch.int64, device=self.device)
        return self._consts[key]

    def _tile_unsigned(self, a, lvl=0, mult_type=-1):
        return self.mont_ops.tile_unsigned(a, [self._two_q(lvl, mult_type)])

    def _enter_ntt(self, a, mult_type=-1):
        self.ntt2_ops.enter_ntt_radix2(a, None, None, None, self._sp(mult_type))

    def _intt_exit(self, a, mult_type=-1):
        self.ntt2_ops.intt_radix2_exit(a, None, None, None, self._sp(mult_type))

    # ---- constants (ckks_engine.py:241-291) --------------------------------------------------------
    def _init_keygen(self, seed=None, nonce=None):
        from .rng import Csprng

        ctx = self.ctx
        # the op-layer wrappers bound to THIS engine's context (not the device-current one)
        self.mont_ops = wrapper.bind(wrapper.mont_ops, ctx)
        self.ntt2_ops = wrapper.bind(wrapper.ntt2_ops, ctx)
        self.he_ops = wrapper.bind(wrapper.he_ops, ctx)
        self._consts = {}
        self.rng = Csprng(num_coefs=self.N, num_channels=[ctx.num_ordinary],
                          num_repeating_channels=max(ctx.K, 2), devices=[str(self.device)], seed=seed, nonce=nonce)
        R = 1 << 62
        Pprod = math.prod(ctx.q[-ctx.K:])
        self.mont_PR = [torch.tensor([(Pprod * R) % ctx.q[i] for i in range(ctx.num_ordinary)], dtype=torch.int64,
                                     device=self.device)]
        base = ctx.q[ctx.num_ordinary - 1]
        self.base_prime = base
        # final_q[level] = first prime alive at that level; final_scalar = q^-1 R mod base
        self.final_scalar = [torch.tensor([(pow(ctx.q[l], -1, base) * R) % base], dtype=torch.int64, device=self.device)
                             for l in range(ctx.num_scales)]
        self._sk = self._pk = self._evk = None
        self._rotk = _RotationKeys(self._create_rotation_key)

    # ---- cached keys (ckks_engine.py:308-407) ------------------------------------------------------
    @property
    def sk(self) -> SecretKey:
        if self._sk is None:
            self._sk = self._create_secret_key()
        return self._sk

    @sk.setter
    def sk(self, sk):
        self._sk, self._pk, self._evk = sk, None, None
        self._rotk.clear()

    @property
    def pk(self) -> PublicKey:
        if self._pk is None:
            self._pk = self._create_public_key()
        return self._pk

    @pk.setter
    def pk(self, pk):
        self._pk = pk

    @property
    def evk(self) -> EvaluationKey:
        if self._evk is None:
            self._evk = self._create_evk()
        return self._evk

    @evk.setter
    def evk(self, evk):
        self._evk = evk

    @property
    def rotk(self):
        return self._rotk

    # ---- secret / public key (ckks_engine.py:486-557) -------------------------------------------------
    def _create_secret_key(self, include_special: bool = True) -> SecretKey:
        ternary = self.rng.randint(amax=3, shift=-1, repeats=1)
        mult_type = -2 if include_special else -1
        s = self._tile_unsigned(ternary, 0, mult_type)
        self._enter_ntt(s, mult_type)
        return SecretKey(data=s, flags=(FLAGS.INCLUDE_SPECIAL if include_special else FLAGS(0))
                         | FLAGS.MONTGOMERY_STATE | FLAGS.NTT_STATE, level=0, logN=self.logN)

    def _create_public_key(self, sk: SecretKey = None, *, include_special: bool = False, a=None) -> PublicKey:
        """pk = (e - a s, a)."""
        sk = sk or self.sk
        if include_special and not sk.has_flag(FLAGS.INCLUDE_SPECIAL):
            raise ValueError("the secret key does not include the special primes")
        mult_type = -2 if include_special else -1
        e = self.rng.discrete_gaussian(repeats=1)
        e = self._tile_unsigned(e, 0, mult_type)
        self._enter_ntt(e, mult_type)
        repeats = self.ctx.K if sk.has_flag(FLAGS.INCLUDE_SPECIAL) else 0
        if a is None:
            a = self.rng.randint([[self.ctx.q[i] for i in self._primes(0, mult_type)]], repeats=repeats)
        sa = self.mont_ops.mont_mult(a, sk.data, self._sp(mult_type))
        pk0 = self.mont_ops.mont_sub(e, sa, self._sp(mult_type))
        return PublicKey(data=[pk0, a], flags=(FLAGS.INCLUDE_SPECIAL if include_special else FLAGS(0))
                         | FLAGS.MONTGOMERY_STATE | FLAGS.NTT_STATE, level=0, logN=self.logN)

    # ---- encrypt / decrypt (ckks_engine.py:565-760) -----------------------------------------------------
    def encrypt(self, pt, pk: PublicKey = None, *, level: int = 0) -> Ciphertext:
        """pt: per-device list with one int64 [N] (or [1, N]) encoded message."""
        pk = pk or self.pk
        mult_type = -2 if pk.has_flag(FLAGS.INCLUDE_SPECIAL) else -1
        sp = self._sp(mult_type)
        e0e1 = self.rng.discrete_gaussian(repeats=2)
        e0, e1 = [e[0] for e in e0e1], [e[1] for e in e0e1]
        e0_t = self._tile_unsigned(e0, level, mult_type)
        e1_t = self._tile_unsigned(e1, level, mult_type)
        pt_t = self._tile_unsigned(pt, level, mult_type)
        self.mont_ops.mont_enter_Rs_scale(pt_t, sp)
        self.mont_ops.mont_reduce(pt_t, sp)
        pte0 = self.mont_ops.mont_add(pt_t, e0_t, sp)
        pk0, pk1 = [pk.data[0][0][level:]], [pk.data[1][0][level:]]
        v = self.rng.randint(amax=2, shift=0, repeats=1)
        v = self._tile_unsigned(v, level, mult_type)
        self._enter_ntt(v, mult_type)
        vpk0 = self.mont_ops.mont_mult(v, pk0, sp)
        vpk1 = self.mont_ops.mont_mult(v, pk1, sp)
        self._intt_exit(vpk0, mult_type)
        self._intt_exit(vpk1, mult_type)
        ct0 = self.mont_ops.mont_add_reduce_2q(vpk0, pte0, sp)
        ct1 = self.mont_ops.mont_add_reduce_2q(vpk1, e1_t, sp)
        # (the reference's `encrypt` tags its output NTT|MONTGOMERY although it is in neither state,
        # ckks_engine.py:621-629; `encodecrypt` :2259-2267 tags it correctly -- followed here)
        return Ciphertext(data=[ct0, ct1], flags=(FLAGS.INCLUDE_SPECIAL if pk.has_flag(FLAGS.INCLUDE_SPECIAL)
                                                   else FLAGS(0)), level=level, logN=self.logN)

    def _final_scale(self, pt, level, include_special, final_round):
        """Shared tail of decrypt_double / decrypt_triplet (ckks_engine.py:677-705, 738-760)."""
        base_at = -self.ctx.K - 1 if include_special else -1
        base = pt[0][base_at][None, :]
        scaler = pt[0][0][None, :]
        scaled = self.mont_ops.mont_sub([base], [scaler], self.ctx.K)
        self.mont_ops.mont_enter_scalar(scaled, [self.final_scalar[level]], self.ctx.K)
        self.mont_ops.reduce_2q(scaled, self.ctx.K)
        self.mont_ops.make_signed(scaled, self.ctx.K)
        if final_round:
            # the reference reads qlists[0][-K-2]: the LAST scale prime at every level (ckks_engine.py:696-703)
            rounding_prime = self.ctx.q[self.ctx.num_ordinary - 2]
            scaled[0] += (scaler[0] > (rounding_prime // 2)) * 1
        return scaled

    def decrypt_double(self, ct: Ciphertext, sk: SecretKey = None, *, final_round=True):
        sk = sk or self.sk
        self._require_plain(ct)
        level = ct.level
        ct0 = ct.data[0][0]
        sk_data = sk.data[0][level:]
        a = ct.data[1][0].clone()
        self._enter_ntt([a])
        sa = self.mont_ops.mont_mult([a], [sk_data], self.ctx.K)
        self._intt_exit(sa)
        pt = self.mont_ops.mont_add_reduce_2q([ct0], sa, self.ctx.K)
        return self._final_scale(pt, level, ct.has_flag(FLAGS.INCLUDE_SPECIAL), final_round)

    def decrypt_triplet(self, ct_mult, sk: SecretKey = None, *, final_round=True):
        sk = sk or self.sk
        level, K = ct_mult.level, self.ctx.K
        d0 = [ct_mult.data[0][0].clone()]
        d1, d2 = [ct_mult.data[1][0]], [ct_mult.data[2][0]]
        self.ntt2_ops.intt_radix2_exit_reduce(d0, None, None, None, K)
        sk_data = [sk.data[0][level:]]
        d1_s = self.mont_ops.mont_mult(d1, sk_data, K)
        s2 = self.mont_ops.mont_mult(sk_data, sk_data, K)
        d2_s2 = self.mont_ops.mont_mult(d2, s2, K)
        self._intt_exit(d1_s)
        self._intt_exit(d2_s2)
        pt = self.mont_ops.mont_add(d0, d1_s, K)
        pt = self.mont_ops.mont_add_reduce_2q(pt, d2_s2, K)
        return self._final_scale(pt, level, ct_mult.has_flag(FLAGS.INCLUDE_SPECIAL), final_round)

    def decrypt(self, ct, sk: SecretKey = None, *, final_round=True):
        from .typing import CiphertextTriplet

        if isinstance(ct, CiphertextTriplet):
            return self.decrypt_triplet(ct, sk, final_round=final_round)
        return self.decrypt_double(ct, sk, final_round=final_round)

    # ---- key-switching keys (ckks_engine.py:796-860, 1621-1634, 1739-1764) -------------------------------
    def _groups(self):
        """global digit groups at level 0: K scale primes each, then the base prime (rns_partition.py:7-52)"""
        ns, K = self.ctx.num_scales, self.ctx.K
        parts = [list(range(i, min(i + K, ns))) for i in range(0, ns, K)]
        parts.append([ns])
        return parts

    def create_key_switching_key(self, sk_from: SecretKey, sk_to: SecretKey, a=None) -> KeySwitchKey:
        for k in (sk_from, sk_to):
            if not (k.has_flag(FLAGS.NTT_STATE) and k.has_flag(FLAGS.MONTGOMERY_STATE)):
                raise ValueError("key-switching keys are built from NTT + Montgomery secret keys")
        Psk = [sk_from.data[0][: self.ctx.num_ordinary].clone()]
        self.mont_ops.mont_enter_scalar(Psk, self.mont_PR, self.ctx.K)
        two_q = self._two_q(0, -2)
        ksk = []
        for gid, part in enumerate(self._groups()):
            crs = a[gid] if a else None
            pk = self._create_public_key(sk_to, include_special=True, a=crs)
            lo, hi = part[0], part[-1] + 1
            pk_rows = pk.data[0][0][lo:hi]
            upd = self.mont_ops.mont_add_legacy([pk_rows], [Psk[0][lo:hi]], [two_q[lo:hi]])[0]
            pk_rows.copy_(upd)
            pk.misc["description"] = f"key switch key part index {gid}"
            ksk.append(pk)
        return KeySwitchKey(data=ksk, flags=FLAGS.INCLUDE_SPECIAL | FLAGS.MONTGOMERY_STATE | FLAGS.NTT_STATE, level=0,
                            logN=self.logN)

    def _create_evk(self, sk: SecretKey = None) -> EvaluationKey:
        sk = sk or self.sk
        sk2 = EvaluationKey(data=self.mont_ops.mont_mult(sk.data, sk.data, 0),
                            flags=FLAGS.MONTGOMERY_STATE | FLAGS.NTT_STATE | FLAGS.INCLUDE_SPECIAL, level=sk.level)
        return EvaluationKey.wrap(self.create_key_switching_key(sk2, sk))

    def _rotate_coefficients(self, m, delta):
        """utils/encoding.py:275-293: out[g n mod N] = (-1)^floor(g n / N) m[n], g = 3^delta mod 2N."""
        N = m.size(-1)
        g = galois_element(N, delta)
        pn = (torch.arange(N, device=m.device, dtype=torch.int64) * g) % (2 * N)
        out = torch.zeros_like(m)
        out[..., pn % N] = torch.where(pn >= N, -m, m)
        return out

    def _create_rotation_key(self, delta: int, a=None, sk: SecretKey = None) -> RotationKey:
        sk = sk or self.sk
        s = [x.clone() for x in sk.data]
        self.ntt2_ops.intt_radix2(s, None, None, None, self.ctx.K)  # NTTContext defaults: mult_type -1
        s = [self._rotate_coefficients(x, delta) for x in s]
        self.ntt2_ops.ntt_radix2(s, None, None, None, self.ctx.K)
        sk_rot = SecretKey(data=s, flags=FLAGS.MONTGOMERY_STATE | FLAGS.NTT_STATE, level=0, logN=self.logN)
        return RotationKey.wrap(self.create_key_switching_key(sk_rot, sk, a=a), delta=delta)


class CodecMixin:
    """encode / decode / encodecrypt / decryptcode / pc_add: ckks_engine.py:258-291
    (deviations, corrections), :437-479, :2178-2422, :2516-2539, with the reference's bias guard (the DC
    coefficient's integral part travels as an exact RNS constant instead of through the float path)."""

    def _init_codec(self, bias_guard=True, norm="forward"):
        import numpy as np

        ctx = self.ctx
        self.bias_guard, self.norm = bias_guard, norm
        self.scale = 2 ** ctx.scale_bits
        self.alpha = [(self.scale / np.float64(q)) ** 2 for q in ctx.q[: ctx.num_scales]]
        self.deviations = [1]
        for al in self.alpha:
            self.deviations.append(self.deviations[-1] ** 2 * al)
        self.final_alpha = [self.scale / np.float64(ctx.q[l]) for l in range(ctx.num_scales)]
        self.corrections = [1 / (d * fa) for d, fa in zip(self.deviations, self.final_alpha)]

    def encode(self, m, level: int = 0, padding=True, scale=None):
        from . import codec

        if padding:
            m = codec.padding(m, num_slots=self.num_slots)
        return [codec.encode(m, scale=scale or self.scale, rng=self.rng, device=str(self.device),
                             deviation=self.deviations[level], norm=self.norm)]

    def decode(self, m, level=0, is_real: bool = False):
        from . import codec

        decoded = codec.decode(m[0].squeeze(), scale=self.scale, correction=self.corrections[level], norm=self.norm)
        out = decoded[: self.N // 2].cpu().numpy()
        return out.real if is_real else out

    def encodecrypt(self, m, pk=None, *, level: int = 0, padding=True) -> Ciphertext:
        import numpy as np

        from . import codec

        pk = pk or self.pk
        if padding:
            m = codec.padding(m=m, num_slots=self.num_slots)
        pt = codec.encode(m, scale=self.scale, device=str(self.device), norm=self.norm, deviation=self.deviations[level],
                          rng=self.rng, return_without_scaling=self.bias_guard)
        mult_type = -2 if pk.has_flag(FLAGS.INCLUDE_SPECIAL) else -1
        dc_rns = None
        if self.bias_guard:
            dc_integral = pt[0].item() // 1
            pt[0] -= dc_integral
            dc_scale = int(dc_integral) * int(self.scale)
            dc_rns = torch.tensor([dc_scale % self.ctx.q[i] for i in self._primes(level, -1)], dtype=torch.int64,
                                  device=self.device)
            pt *= np.float64(self.scale)
            pt = self.rng.randround(pt)
        return self._encrypt_encoded([pt], pk, level, mult_type, dc_rns)

    def _encrypt_encoded(self, encoded, pk, level, mult_type, dc_rns):
        sp = self._sp(mult_type)
        e0e1 = self.rng.discrete_gaussian(repeats=2)
        e0, e1 = [e[0] for e in e0e1], [e[1] for e in e0e1]
        e0_t = self._tile_unsigned(e0, level, mult_type)
        e1_t = self._tile_unsigned(e1, level, mult_type)
        pt_t = self._tile_unsigned(encoded, level, mult_type)
        if dc_rns is not None:
            pt_t[0][:, 0] += dc_rns
        self.mont_ops.mont_enter_Rs_scale(pt_t, sp)
        self.mont_ops.mont_reduce(pt_t, sp)
        pte0 = self.mont_ops.mont_add(pt_t, e0_t, sp)
        pk0, pk1 = [pk.data[0][0][level:]], [pk.data[1][0][level:]]
        v = self.rng.randint(amax=2, shift=0, repeats=1)
        v = self._tile_unsigned(v, level, mult_type)
        self._enter_ntt(v, mult_type)
        vpk0 = self.mont_ops.mont_mult(v, pk0, sp)
        vpk1 = self.mont_ops.mont_mult(v, pk1, sp)
        self._intt_exit(vpk0, mult_type)
        self._intt_exit(vpk1, mult_type)
        ct0 = self.mont_ops.mont_add_reduce_2q(vpk0, pte0, sp)
        ct1 = self.mont_ops.mont_add_reduce_2q(vpk1, e1_t, sp)
        return Ciphertext(data=[ct0, ct1], flags=(FLAGS.INCLUDE_SPECIAL if pk.has_flag(FLAGS.INCLUDE_SPECIAL)
                                                   else FLAGS(0)), level=level, logN=self.logN)

    def decryptcode(self, ct, sk=None, *, is_real=False, final_round=True):
        from . import codec
        from .typing import CiphertextTriplet

        sk = sk or self.sk
        level, K = ct.level, self.ctx.K
        sk_data = sk.data[0][level:]
        if isinstance(ct, CiphertextTriplet):
            d0 = [ct.data[0][0].clone()]
            self.ntt2_ops.intt_radix2_exit_reduce(d0, None, None, None, K)
            d1_s = self.mont_ops.mont_mult([ct.data[1][0]], [sk_data], K)
            s2 = self.mont_ops.mont_mult([sk_data], [sk_data], K)
            d2_s2 = self.mont_ops.mont_mult([ct.data[2][0]], s2, K)
            self._intt_exit(d1_s)
            self._intt_exit(d2_s2)
            pt = self.mont_ops.mont_add(d0, d1_s, K)
            pt = self.mont_ops.mont_add(pt, d2_s2, K)
            self.mont_ops.reduce_2q(pt, K)
        else:
            self._require_plain(ct)
            a = ct.data[1][0].clone()
            self._enter_ntt([a])
            sa = self.mont_ops.mont_mult([a], [sk_data], K)
            self._intt_exit(sa)
            pt = self.mont_ops.mont_add([ct.data[0][0]], sa, K)
            self.mont_ops.reduce_2q(pt, K)
        include_special = ct.has_flag(FLAGS.INCLUDE_SPECIAL)
        base_at = -K - 1 if include_special else -1
        base = pt[0][base_at][None, :]
        scaler = pt[0][0][None, :]
        alive = self._primes(level, -1)
        guard = len(alive) >= 3 and self.bias_guard
        dc = 0
        if guard:  # exact CRT of the DC coefficient over (base, first, second) primes (ckks_engine.py:2352-2383)
            dc0, dc1, dc2 = base[0][0].item(), scaler[0][0].item(), pt[0][1][0].item()
            base[0][0] = 0
            scaler[0][0] = 0
            q0 = self.base_prime
            q1, q2 = self.ctx.q[alive[0]], self.ctx.q[alive[1]]
            Q = q0 * q1 * q2
            Q0, Q1, Q2 = q1 * q2, q0 * q2, q0 * q1
            dc = (dc0 * pow(Q0, -1, q0) * Q0 + dc1 * pow(Q1, -1, q1) * Q1 + dc2 * pow(Q2, -1, q2) * Q2) % Q
            dc = dc if dc <= Q // 2 else dc - Q
            dc = (dc + (q1 - 1)) // q1
        scaled = self.mont_ops.mont_sub([base], [scaler], K)
        self.mont_ops.mont_enter_scalar(scaled, [self.final_scalar[level]], K)
        self.mont_ops.reduce_2q(scaled, K)
        self.mont_ops.make_signed(scaled, K)
        if final_round:
            rounding_prime = self.ctx.q[self.ctx.num_ordinary - 2]
            scaled[0] += (scaler[0] > (rounding_prime // 2)) * 1
        correction = self.corrections[level]
        decoded = codec.decode(scaled[0][-1], scale=self.scale, correction=correction, norm=self.norm,
                               return_without_scaling=self.bias_guard)
        decoded = decoded[: self.N // 2].cpu().numpy()
        # (as in the reference, the scaling is applied here -- a second time when bias_guard is off)
        decoded = decoded / self.scale * correction
        if guard:
            decoded += dc / self.scale * correction
        return decoded.real if is_real else decoded

    # ---- plaintext operands (ckks_engine.py:2516-2560) ---------------------------------------------------
    def _plain_operand(self, pt, level, kind):
        cache = pt.cache[level]
        if kind not in cache:
            m = pt.src * math.sqrt(self.deviations[level + 1])
            p = self.encode(m, level, scale=pt.scale)
            p = self._tile_unsigned(p, level)
            if kind == "pc_add":
                self.mont_ops.mont_enter_Rs_scale(p, self.ctx.K)
            else:
                self._enter_ntt(p)
            cache[kind] = p
        return cache[kind]

    def pc_add(self, pt, ct: Ciphertext, inplace: bool = False) -> Ciphertext:

        p = self._plain_operand(pt, ct.level, "pc_add")
        new_d0 = self.he_ops.pc_add_fused(ct.data[0], p, self.ctx.K)
        if inplace:
            ct.data[0] = new_d0
            return ct
        return Ciphertext(data=[new_d0, [d.clone() for d in ct.data[1]]], flags=ct._flags, level=ct.level,
                          misc=dict(ct.misc))


def decompose_with_power_of_2(a: int, n: int) -> list:
    """Binary expansion of the offset a (mod n) (tiberate/utils/massive.py:83-98)."""
    if n <= 0 or n & (n - 1):
        raise AssertionError("n must be a power of 2")
    a = a + n if a < 0 else a
    return [1 << e for e in range(n.bit_length() - 1) if a & (1 << e)]


def decompose_rot_offsets(offset: int, num_slots: int, rotks) -> list:
    """Fewest-steps decomposition of `offset` over the rotation keys already made plus the powers of two
    below num_slots / 2, never longer than the binary expansion (utils/massive.py:101-146): breadth-first
    over partial sums in [-num_slots, num_slots], candidate steps in increasing order."""
    from collections import deque

    best = decompose_with_power_of_2(offset, num_slots)
    steps = sorted(set(list(rotks.keys()) + [1 << i for i in range(int(math.log2(num_slots // 2)))]))
    seen, queue = {0}, deque([(0, [])])
    while queue:
        total, path = queue.popleft()
        if total == offset:
            if len(path) <= len(best):
                return path
            break
        for s in steps:
            nxt = total + s
            if -num_slots <= nxt <= num_slots and nxt not in seen:
                seen.add(nxt)
                queue.append((nxt, [*path, s]))
    return best


class LevelMixin:
    """Level management and multi-step rotation: ckks_engine.py:1908-1926
    (rotate_offset), :2086-2171 (level_up), :2473-2482 (negate)."""

    def rotate_offset(self, ct: Ciphertext, offset: int, inplace: bool = True, return_decomposed_offsets=False):
        if offset == 0:
            return ct if inplace else ct.clone()
        if offset in self.rotk:
            return self.rotate_single(ct, self.rotk[offset])
        offsets = decompose_rot_offsets(offset, self.num_slots, rotks=self.rotk)
        for delta in offsets:
            ct = self.rotate_single(ct, self.rotk[delta])
        return (ct, offsets) if return_decomposed_offsets else ct

    def level_up(self, ct: Ciphertext, dst_level: int, inplace=False) -> Ciphertext:
        """Rescale once, drop to dst_level and multiply by round(scale * deviation ratio) so that the
        message keeps the scale the destination level expects."""
        import numpy as np

        if ct.level == dst_level:
            return ct if inplace else ct.clone()
        new_ct = self.rescale(ct)
        src_level = ct.level + 1
        deviated_delta = round(self.scale * (self.deviations[dst_level] / np.sqrt(self.deviations[src_level])))
        drop = dst_level - src_level
        d0 = [new_ct.data[0][0][drop:]] if drop > 0 else new_ct.data[0]
        d1 = [new_ct.data[1][0][drop:]] if drop > 0 else new_ct.data[1]
        mult = [torch.tensor([(deviated_delta * (1 << 62)) % self.ctx.q[i] for i in self._primes(dst_level, -1)],
                             dtype=torch.int64, device=self.device)]
        K = self.ctx.K
        self.mont_ops.mont_enter_scalar(d0, mult, K)
        self.mont_ops.mont_enter_scalar(d1, mult, K)
        self.mont_ops.reduce_2q(d0, K)
        self.mont_ops.reduce_2q(d1, K)
        return Ciphertext(data=[d0, d1], level=dst_level, logN=self.logN, misc=dict(new_ct.misc))

    def negate(self, ct: Ciphertext, inplace: bool = False) -> Ciphertext:
        if not inplace:
            ct = ct.clone()
        for part in ct.data:
            for d in part:
                d *= -1
            self.mont_ops.make_signed(part, self.ctx.K)
        return ct
