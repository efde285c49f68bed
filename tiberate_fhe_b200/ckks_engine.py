"""CkksEngine hot-path mirror on the B200 backend.

Same method names, argument meaning and error behaviour as tiberate/ckks_engine.py for the hot
methods: rescale :1520, cc_mult :1640, relinearize :1695, create_switcher :1201,
switch_key :1403, rotate_single :1804, cc_add_double :1932 / cc_sub_double :2011, pc_mult :2542.
Every call is ONE fused C-ABI entry (include/tb200.h engine layer) instead of the reference's
hundreds of op launches; outputs are bit-identical (tests/test_gpu_parity.py, tests/golden_check.py).

Differences, all deliberate:
  * inputs are never modified (the reference's cc_mult(pre_rescale=False) / relinearize transform
    their arguments in place); rescale returns fresh tensors instead of storage-offset views;
  * tensors may carry a leading batch dimension ([B, limbs, N]) -- a batch of ciphertexts sharing
    keys is processed by one call (BASELINE.json configs[2]);
  * single device per engine: limb sharding across devices (reference `devices=[...]`) is replaced
    by batch sharding across ranks (tiberate_fhe_b200/dist.py).
Key generation, encryption and decryption live in keygen.py, the CSPRNG (8f-1) in
rng/csprng.py; `seed` (8 words) / `nonce` (2 words) fix the CSPRNG stream, None draws from os.urandom.
"""

from __future__ import annotations

import torch

from . import wrapper
from .context import KeySwitchKeyView, Tb200Context, galois_element
from .keygen import CodecMixin, KeyGenMixin, LevelMixin
from .presets import PRESETS
from .typing import FLAGS, Ciphertext, CiphertextTriplet, KeySwitchKey, Plaintext


class MaximumLevelError(Exception):
    """tiberate/errors.py: raised when an operation would go past the last level."""

    def __init__(self, level, level_max):
        super().__init__(f"Cannot go below the maximum level: level={level}, level_max={level_max}")


class NTTStateError(Exception):
    def __init__(self, expected):
        super().__init__(f"NTT state mismatch: expected NTT_STATE={expected}")


class MontgomeryStateError(Exception):
    def __init__(self, expected):
        super().__init__(f"Montgomery state mismatch: expected MONTGOMERY_STATE={expected}")


class CkksEngine(KeyGenMixin, CodecMixin, LevelMixin):
    def __init__(self, ckks_config=None, devices=None, *, chunk: int = 4, seed=None, nonce=None, bias_guard=True,
                 norm="forward"):
        """ckks_config: None (reference default: logN15 preset), an int logN naming a preset, or a dict
        with keys logN, q (prime chain [scale..., base, special...]), num_special_primes[, scale_bits]."""
        if ckks_config is None:
            ckks_config = 15  # ckks_engine.py:52-54
        if isinstance(ckks_config, int):
            p = PRESETS[ckks_config]
            ckks_config = dict(logN=ckks_config, q=p["q"], num_special_primes=p["K"])
        devices = devices or ["cuda:0"]
        if len(devices) != 1:
            raise ValueError("one device per engine: shard the ciphertext batch across ranks (dist.py)")
        self.device = torch.device(devices[0])
        idx = self.device.index if self.device.index is not None else 0
        self.ctx = Tb200Context(ckks_config["logN"], ckks_config["q"], ckks_config["num_special_primes"],
                                ckks_config.get("scale_bits", 40), device=idx)
        self.ctx.set_chunk(chunk)
        wrapper.set_context(self.ctx)
        self.logN, self.N = self.ctx.logN, self.ctx.N
        self.num_special_primes = self.ctx.K
        self._init_keygen(seed, nonce)  # CSPRNG + key-generation constants (keygen.py)
        self._init_codec(bias_guard, norm)

    @property
    def num_levels(self) -> int:  # ckks_engine.py:102-104
        return self.ctx.num_scales

    @property
    def num_slots(self) -> int:
        return self.N // 2

    # ---- helpers -------------------------------------------------------------------------------
    def _key(self, ksk: KeySwitchKey) -> KeySwitchKeyView:
        """KeySwitchKey.data = list over global digit-group id of PublicKey(data=[[b],[a]])."""
        v = getattr(ksk, "_tb200_view", None)  # cached on the key object: lives and dies with it
        if v is None:
            parts = []
            for part in ksk.data:
                if part is None or (isinstance(part, list) and len(part) == 0):
                    parts.append(None)
                else:
                    parts.append((part.data[0][0], part.data[1][0]))
            v = KeySwitchKeyView(parts, self.N)
            ksk._tb200_view = v
        return v

    @staticmethod
    def _t(poly_list):
        return poly_list[0]

    def _empty_like_rows(self, ref, rows):
        shape = list(ref.shape)
        shape[-2] = rows
        return torch.empty(shape, dtype=torch.int64, device=ref.device)

    @staticmethod
    def _require_plain(ct):
        if ct.has_flag(FLAGS.NTT_STATE):
            raise NTTStateError(expected=False)
        if ct.has_flag(FLAGS.MONTGOMERY_STATE):
            raise MontgomeryStateError(expected=False)

    # ---- rescale -------------------------------------------------------------------------------
    def rescale(self, ct: Ciphertext, exact_rounding=True, inplace=False) -> Ciphertext:
        level = ct.level
        if level + 1 >= self.num_levels:
            raise MaximumLevelError(level=ct.level, level_max=self.num_levels)
        c0, c1 = self._t(ct.data[0]), self._t(ct.data[1])
        L = c0.shape[-2] - 1
        o0, o1 = self._empty_like_rows(c0, L), self._empty_like_rows(c1, L)
        self.ctx.rescale(level, c0, c1, o0, o1, bool(exact_rounding))
        return Ciphertext(data=[[o0], [o1]], level=level + 1, misc=dict(ct.misc))

    # ---- multiplication ------------------------------------------------------------------------
    def cc_mult(self, a: Ciphertext, b: Ciphertext, evk=None, *, pre_rescale=True, post_relin=True):
        evk = evk or (self.evk if post_relin else None)
        level = a.level
        if pre_rescale and level + 1 >= self.num_levels:
            raise MaximumLevelError(level=level, level_max=self.num_levels)
        a0, a1, b0, b1 = self._t(a.data[0]), self._t(a.data[1]), self._t(b.data[0]), self._t(b.data[1])
        rows = a0.shape[-2] - (1 if pre_rescale else 0)
        lvl = level + (1 if pre_rescale else 0)
        if post_relin:
            o0, o1 = self._empty_like_rows(a0, rows), self._empty_like_rows(a0, rows)
            self.ctx.cc_mult_relin(level, a0, a1, b0, b1, self._key(evk), o0, o1, bool(pre_rescale))
            return Ciphertext(data=[[o0], [o1]], level=lvl, misc=dict(a.misc))
        d = [self._empty_like_rows(a0, rows) for _ in range(3)]
        self.ctx.cc_mult_triplet(level, a0, a1, b0, b1, d[0], d[1], d[2], bool(pre_rescale))
        return CiphertextTriplet(data=[[d[0]], [d[1]], [d[2]]],
                                 flags=FLAGS.NTT_STATE | FLAGS.MONTGOMERY_STATE | FLAGS.NEED_RELINERIZE,
                                 level=lvl, misc=dict(a.misc))

    def relinearize(self, ct_triplet: CiphertextTriplet, evk=None) -> Ciphertext:
        evk = evk or self.evk
        if not ct_triplet.has_flag(FLAGS.NTT_STATE):
            raise NTTStateError(expected=True)
        if not ct_triplet.has_flag(FLAGS.MONTGOMERY_STATE):
            raise MontgomeryStateError(expected=True)
        d0, d1, d2 = (self._t(x) for x in ct_triplet.data)
        o0, o1 = torch.empty_like(d0), torch.empty_like(d1)
        self.ctx.relinearize(ct_triplet.level, d0, d1, d2, self._key(evk), o0, o1)
        return Ciphertext(data=[[o0], [o1]], level=ct_triplet.level, misc=dict(ct_triplet.misc))

    # ---- key switching -------------------------------------------------------------------------
    def create_switcher(self, a, ksk: KeySwitchKey, level: int, exit_ntt: bool = False):
        """a: per-device list with one [L, N] coefficient-domain canonical polynomial."""
        x = self._t(a)
        if exit_ntt:
            x = x.clone()
            self.ctx.intt(x, level, 2)
        o0, o1 = torch.empty_like(x), torch.empty_like(x)
        self.ctx.keyswitch(level, x, self._key(ksk), o0, o1)
        return [o0], [o1]

    def switch_key(self, ct: Ciphertext, ksk: KeySwitchKey) -> Ciphertext:
        c0, c1 = self._t(ct.data[0]), self._t(ct.data[1])
        if ct.has_flag(FLAGS.NTT_STATE):
            c1 = c1.clone()
            self.ctx.intt(c1, ct.level, 2)
        o0, o1 = torch.empty_like(c0), torch.empty_like(c1)
        self.ctx.switch_key(ct.level, c0, c1, self._key(ksk), o0, o1)
        return Ciphertext(data=[[o0], [o1]], flags=ct._flags, level=ct.level, misc=dict(ct.misc))

    # ---- rotation ------------------------------------------------------------------------------
    def rotate_single(self, ct: Ciphertext, rotk, post_key_switching=True) -> Ciphertext:
        c0, c1 = self._t(ct.data[0]), self._t(ct.data[1])
        g = galois_element(self.N, rotk.delta)
        o0, o1 = torch.empty_like(c0), torch.empty_like(c1)
        self.ctx.rotate(ct.level, g, c0, c1, self._key(rotk) if post_key_switching else None, o0, o1)
        return Ciphertext(data=[[o0], [o1]], flags=ct._flags, level=ct.level, misc=dict(ct.misc))

    def rotate_hoisted(self, ct: Ciphertext, deltas, rotks=None):
        """Rotations of ONE ciphertext by every delta in `deltas` (keys self.rotk[delta] unless given): the ModUp
        digits, extension and forward transform are computed once for all of them (tb200_rotate_hoisted).
        An extension beyond the reference, whose rotate_offset / sum / BSGS loops call rotate_single per key
        (ckks_engine.py:1908-1926, 2770-2788); each result decrypts like rotate_single(ct, rotk[delta])."""
        deltas = list(deltas)
        keys = [(rotks[d] if rotks is not None else self.rotk[d]) for d in deltas]
        c0, c1 = self._t(ct.data[0]), self._t(ct.data[1])
        o0 = torch.empty((len(deltas), *c0.shape), dtype=torch.int64, device=c0.device)
        o1 = torch.empty_like(o0)
        self.ctx.rotate_hoisted(ct.level, [galois_element(self.N, k.delta) for k in keys], c0, c1,
                                [self._key(k) for k in keys], o0, o1)
        return [Ciphertext(data=[[o0[r]], [o1[r]]], flags=ct._flags, level=ct.level, misc=dict(ct.misc))
                for r in range(len(deltas))]

    # ---- add / sub -----------------------------------------------------------------------------
    def _addsub(self, a, b, sub):
        self._require_plain(a)
        self._require_plain(b)
        a0, a1, b0, b1 = self._t(a.data[0]), self._t(a.data[1]), self._t(b.data[0]), self._t(b.data[1])
        o0, o1 = torch.empty_like(a0), torch.empty_like(a1)
        self.ctx.cc_addsub(a.level, sub, a0, a1, b0, b1, o0, o1)
        return Ciphertext(data=[[o0], [o1]], level=a.level, misc=dict(a.misc))

    def cc_add_double(self, a: Ciphertext, b: Ciphertext) -> Ciphertext:
        return self._addsub(a, b, False)

    def cc_sub_double(self, a: Ciphertext, b: Ciphertext) -> Ciphertext:
        return self._addsub(a, b, True)

    cc_add = cc_add_double
    cc_sub = cc_sub_double

    # ---- plaintext multiplication ----------------------------------------------------------------
    def pc_mult(self, pt: Plaintext, ct: Ciphertext, inplace: bool = False, post_rescale=True) -> Ciphertext:
        """pt.cache[level]['pc_mult'] must hold the NTT+Montgomery plaintext [L, N]
        (the reference fills it on first use by encode -> tile_unsigned -> enter_ntt_radix2)."""
        level = ct.level
        cache = pt.cache[level]
        if "pc_mult" not in cache:
            if getattr(pt, "src", None) is None:
                raise KeyError("plaintext has neither a message nor an NTT-form cache for this level")
            self._plain_operand(pt, level, "pc_mult")  # encode -> tile_unsigned -> enter_ntt_radix2
        if post_rescale and level + 1 >= self.num_levels:
            raise MaximumLevelError(level=level, level_max=self.num_levels)
        c0, c1 = self._t(ct.data[0]), self._t(ct.data[1])
        rows = c0.shape[-2] - (1 if post_rescale else 0)
        o0, o1 = self._empty_like_rows(c0, rows), self._empty_like_rows(c1, rows)
        self.ctx.pc_mult(level, self._t(cache["pc_mult"]), c0, c1, o0, o1, bool(post_rescale))
        return Ciphertext(data=[[o0], [o1]], level=level + (1 if post_rescale else 0), misc=dict(ct.misc))
