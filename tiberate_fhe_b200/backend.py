"""install_as_tiberate_backend(): make an UNMODIFIED tiberate (reference) engine run its hot-path
operators on libtb200.

The reference reaches its torch ops only through the module-level functions of
tiberate.libs.wrapper.{mont_ops,ntt2_ops,he_ops}, looked up as module attributes at call time
(tiberate/context/ntt_context.py:11-14,715-871; tiberate/ckks_engine.py:19), so re-pointing those
attributes swaps the backend without touching the engine -- every operator its Python calls, the CSPRNG
ones and the constant pool included.  The reference's own extension stays loaded, which allows
same-process A/B comparison: `uninstall()` restores the original functions.
"""

from __future__ import annotations

import importlib

from . import wrapper
from .context import Tb200Context

_HOT = {
    "mont_ops": ["mont_mult", "mont_add", "mont_sub", "mont_add_reduce_2q", "mont_sub_reduce_2q",
                 "mont_enter_scalar", "mont_enter_Rs", "mont_enter_Rs_scale", "mont_reduce", "reduce_2q",
                 "make_signed", "make_unsigned", "mont_enter", "mont_add_legacy", "tile_unsigned",
                 "mont_add_many_3d", "mont_reduce_add_many_3d", "mont_enter_scalar_reduce_2q"],
    "ntt2_ops": ["ntt_radix2", "enter_ntt_radix2", "intt_radix2", "intt_radix2_exit",
                 "intt_radix2_exit_reduce", "intt_radix2_exit_reduce_signed"],
    "he_ops": ["pc_add_fused", "rescale_exact_rounding_fused", "rescale_non_exact_rounding_fused",
               "switch_key_switch_later_part_extend", "codec_rotate_make_unsigned_reduce_2q",
               "create_switcher_divide_by_p"],
    # SURVEY 8f-1 and the constant pool: with these the reference's Python runs without calling its own
    # extension at all
    "csprng_ops": ["chacha20", "randint_fast", "discrete_gaussian_fast", "randint", "discrete_gaussian", "randround"],
    "const_pool": ["upload_tensor_list", "read_constant_chunk"],
}
_saved = {}


def install_as_tiberate_backend(engine, fused: bool = False) -> list[Tb200Context]:
    """`engine` is a constructed reference CkksEngine; one Tb200Context per engine device is created
    from that device's local prime list (rnsPart.d_special) and made current.

    fused=False: only the operators are re-pointed -- the reference's Python still sequences ~520 op calls per
    HMult, now served by libtb200.  fused=True (single-device engines): additionally the engine's hot METHODS
    (rescale, cc_mult, relinearize, create_switcher, switch_key, rotate_single, pc_mult, cc_add_double,
    cc_sub_double -- tiberate/ckks_engine.py:1520,1640,1695,1201,1403,1804,2542,1932,2011) are bound to
    the one-call-per-method C entries; they take and return the reference's own Ciphertext / CiphertextTriplet
    objects with the same flags, level and metadata, and bit-identical tensors.  Calls outside the fused
    domain (ciphertexts carrying the special limbs, inplace=True on pc_mult, pre_rescale=False) go to the
    reference's own method body, i.e. the op-level backend."""
    import tiberate.libs.wrapper as ref_wrapper

    cfg = engine.ckksCfg
    ctxs = []
    for dev_id, dev in enumerate(engine.nttCtx.devices):
        local = [int(cfg.q[i]) for i in engine.rnsPart.d_special[dev_id]]
        import torch

        idx = torch.device(dev).index or 0
        ctx = Tb200Context(cfg.logN, local, cfg.num_special_primes, cfg.scale_bits, device=idx)
        wrapper.set_context(ctx)
        ctxs.append(ctx)
    for mod_name, names in _HOT.items():
        ref_mod = importlib.import_module(f"tiberate.libs.wrapper.{mod_name}")
        ours = getattr(wrapper, mod_name)
        for n in names:
            _saved.setdefault((mod_name, n), getattr(ref_mod, n))
            setattr(ref_mod, n, getattr(ours, n))
    if fused:
        if len(ctxs) != 1:
            raise ValueError("fused=True needs a single-device engine")
        _install_fused(engine, ctxs[0])
    return ctxs


_FUSED_METHODS = ("rescale", "cc_mult", "relinearize", "create_switcher", "switch_key", "rotate_single", "pc_mult",
                  "cc_add_double", "cc_sub_double")
_fused_engines = []


def _install_fused(engine, ctx: Tb200Context) -> None:
    """Bind the fused C entries as instance attributes of a reference engine (its other methods reach them
    through `self.<name>`, e.g. cc_add -> cc_add_double, rotate_offset -> rotate_single)."""
    import types

    import torch
    from tiberate import errors
    from tiberate.typing import FLAGS, Ciphertext, CiphertextTriplet

    from .context import KeySwitchKeyView, galois_element

    orig = {n: getattr(engine, n) for n in _FUSED_METHODS}  # bound methods of the reference class
    K = engine.ckksCfg.num_special_primes
    N = ctx.N

    def meta(self, misc):
        return dict(logN=self.ckksCfg.logN, creator_hash=self.hash, misc=misc)

    def keyview(ksk) -> KeySwitchKeyView:
        v = getattr(ksk, "_tb200_view", None)
        if v is None:  # KeySwitchKey.data: per digit group a PublicKey with data [[b], [a]] ([P, N] each)
            v = KeySwitchKeyView([None if p is None else (p.data[0][0], p.data[1][0]) for p in ksk.data], N)
            try:
                ksk._tb200_view = v  # lives and dies with the key object
            except AttributeError:
                pass
        return v

    def plain(ct) -> bool:
        """coefficient-domain ciphertext without the special limbs, single device"""
        return (not ct.has_flag(FLAGS.INCLUDE_SPECIAL) and not ct.has_flag(FLAGS.NTT_STATE)
                and not ct.has_flag(FLAGS.MONTGOMERY_STATE) and len(ct.data[0]) == 1)

    def rows_out(like, rows, offset_view):
        """[rows, N] output; offset_view: as rows 1.. of a [rows + 1, N] buffer -- the reference returns such
        storage-offset views from rescale (ckks_engine.py:1553-1554: data[1:])."""
        if offset_view:
            buf = torch.empty((rows + 1, N), dtype=torch.int64, device=like.device)
            return buf, buf[1:]
        t = torch.empty((rows, N), dtype=torch.int64, device=like.device)
        return t, t

    def rescale(self, ct, exact_rounding=True, inplace=False):
        if not plain(ct):
            return orig["rescale"](ct, exact_rounding, inplace)
        level = ct.level
        if level + 1 >= self.num_levels:
            raise errors.MaximumLevelError(level=ct.level, level_max=self.num_levels)
        c0, c1 = ct.data[0][0], ct.data[1][0]
        L = c0.shape[0] - 1
        if inplace:  # the reference rescales rows 1.. of the ciphertext's own tensors and returns views of them
            o0, o1 = c0[1:], c1[1:]
        else:  # ... or of a clone: row 0 of the underlying buffer is the dropped limb
            b0, o0 = rows_out(c0, L, True)
            b1, o1 = rows_out(c1, L, True)
            b0[0].copy_(c0[0])
            b1[0].copy_(c1[0])
        ctx.rescale(level, c0, c1, o0, o1, bool(exact_rounding))
        return Ciphertext(data=[[o0], [o1]], level=level + 1, **meta(self, ct.misc))

    def cc_mult(self, a, b, evk=None, *, pre_rescale=True, post_relin=True):
        if not (pre_rescale and plain(a) and plain(b)):
            return orig["cc_mult"](a, b, evk, pre_rescale=pre_rescale, post_relin=post_relin)
        level = a.level
        if level + 1 >= self.num_levels:
            raise errors.MaximumLevelError(level=a.level, level_max=self.num_levels)
        a0, a1, b0, b1 = a.data[0][0], a.data[1][0], b.data[0][0], b.data[1][0]
        rows = a0.shape[0] - 1
        if post_relin:
            evk = evk or self.evk
            o0, o1 = rows_out(a0, rows, False)[1], rows_out(a0, rows, False)[1]
            ctx.cc_mult_relin(level, a0, a1, b0, b1, keyview(evk), o0, o1, True)
            return Ciphertext(data=[[o0], [o1]], level=level + 1, **meta(self, a.misc))
        d = [rows_out(a0, rows, False)[1] for _ in range(3)]
        ctx.cc_mult_triplet(level, a0, a1, b0, b1, d[0], d[1], d[2], True)
        return CiphertextTriplet(data=[[d[0]], [d[1]], [d[2]]],
                                 flags=FLAGS.NTT_STATE | FLAGS.MONTGOMERY_STATE | FLAGS.NEED_RELINERIZE,
                                 level=level + 1, **meta(self, a.misc))

    def relinearize(self, ct_triplet, evk=None):
        evk = evk or self.evk
        if not ct_triplet.has_flag(FLAGS.NTT_STATE):
            raise errors.NTTStateError(expected=True)
        if not ct_triplet.has_flag(FLAGS.MONTGOMERY_STATE):
            raise errors.MontgomeryStateError(expected=True)
        if len(ct_triplet.data[0]) != 1:
            return orig["relinearize"](ct_triplet, evk)
        d0, d1, d2 = (x[0] for x in ct_triplet.data)
        o0, o1 = torch.empty_like(d0), torch.empty_like(d1)
        # (the reference also leaves its triplet inverse-transformed in place, ckks_engine.py:1709-1711;
        # the fused call does not touch its inputs)
        ctx.relinearize(ct_triplet.level, d0, d1, d2, keyview(evk), o0, o1)
        return Ciphertext(data=[[o0], [o1]], level=ct_triplet.level, **meta(self, ct_triplet.misc))

    def create_switcher(self, a, ksk, level, exit_ntt=False):
        if len(a) != 1 or a[0].shape[0] != ctx.num_ordinary - level:
            return orig["create_switcher"](a, ksk, level, exit_ntt)
        x = a[0]
        if exit_ntt:  # the reference transforms the caller's tensor in place (ckks_engine.py:1236-1237)
            ctx.intt(x, level, 2)
        o0, o1 = torch.empty_like(x), torch.empty_like(x)
        ctx.keyswitch(level, x, keyview(ksk), o0, o1)
        return [o0], [o1]

    def switch_key(self, ct, ksk):
        if ct.has_flag(FLAGS.INCLUDE_SPECIAL) or ct.has_flag(FLAGS.NTT_STATE) or len(ct.data[0]) != 1:
            return orig["switch_key"](ct, ksk)
        c0, c1 = ct.data[0][0], ct.data[1][0]
        o0, o1 = torch.empty_like(c0), torch.empty_like(c1)
        ctx.switch_key(ct.level, c0, c1, keyview(ksk), o0, o1)
        return Ciphertext(data=[[o0], [o1]], flags=ct._flags, level=ct.level, **meta(self, ct.misc))

    def rotate_single(self, ct, rotk, post_key_switching=True):
        if ct.has_flag(FLAGS.INCLUDE_SPECIAL) or ct.has_flag(FLAGS.NTT_STATE) or len(ct.data[0]) != 1:
            return orig["rotate_single"](ct, rotk, post_key_switching)
        c0, c1 = ct.data[0][0], ct.data[1][0]
        o0, o1 = torch.empty_like(c0), torch.empty_like(c1)
        ctx.rotate(ct.level, galois_element(N, rotk.delta), c0, c1, keyview(rotk) if post_key_switching else None,
                   o0, o1)
        return Ciphertext(data=[[o0], [o1]], flags=ct._flags, level=ct.level, **meta(self, ct.misc))

    def pc_mult(self, pt, ct, inplace=False, post_rescale=True):
        if inplace or not plain(ct):
            return orig["pc_mult"](pt, ct, inplace, post_rescale)
        level = ct.level
        key = str(orig["pc_mult"])  # the cache slot the reference's own method uses (ckks_engine.py:2550)
        if key not in pt.cache[level]:
            import math

            m = pt.src * math.sqrt(self.deviations[level + 1])
            pt_ = self.encode(m, level, scale=pt.scale)
            pt_ = self.nttCtx.tile_unsigned(pt_, level)
            self.nttCtx.enter_ntt_radix2(pt_, level)
            pt.cache[level][key] = pt_
        pt_ = pt.cache[level][key]
        if post_rescale and level + 1 >= self.num_levels:
            raise errors.MaximumLevelError(level=level, level_max=self.num_levels)
        c0, c1 = ct.data[0][0], ct.data[1][0]
        rows = c0.shape[0] - (1 if post_rescale else 0)
        o0 = rows_out(c0, rows, post_rescale)[1]
        o1 = rows_out(c1, rows, post_rescale)[1]
        ctx.pc_mult(level, pt_[0], c0, c1, o0, o1, bool(post_rescale))
        return Ciphertext(data=[[o0], [o1]], level=level + (1 if post_rescale else 0), **meta(self, ct.misc))

    def addsub(sub):
        name = "cc_sub_double" if sub else "cc_add_double"

        def fn(self, a, b):
            for x in (a, b):
                if x.has_flag(FLAGS.NTT_STATE):
                    raise errors.NTTStateError(expected=False)
                if x.has_flag(FLAGS.MONTGOMERY_STATE):
                    raise errors.MontgomeryStateError(expected=False)
            if a.has_flag(FLAGS.INCLUDE_SPECIAL) or len(a.data[0]) != 1:
                return orig[name](a, b)
            a0, a1, b0, b1 = a.data[0][0], a.data[1][0], b.data[0][0], b.data[1][0]
            o0, o1 = torch.empty_like(a0), torch.empty_like(a1)
            ctx.cc_addsub(a.level, sub, a0, a1, b0, b1, o0, o1)
            return Ciphertext(data=[[o0], [o1]], level=a.level, **meta(self, a.misc))

        fn.__name__ = name
        return fn

    fused = dict(rescale=rescale, cc_mult=cc_mult, relinearize=relinearize, create_switcher=create_switcher,
                 switch_key=switch_key, rotate_single=rotate_single, pc_mult=pc_mult, cc_add_double=addsub(False),
                 cc_sub_double=addsub(True))
    for n, f in fused.items():
        setattr(engine, n, types.MethodType(f, engine))
    _fused_engines.append(engine)


def uninstall() -> None:
    import tiberate.libs.wrapper as ref_wrapper

    for (mod_name, n), fn in _saved.items():
        setattr(importlib.import_module(f"tiberate.libs.wrapper.{mod_name}"), n, fn)
    _saved.clear()
    for engine in _fused_engines:  # instance attributes shadowing the class methods
        for n in _FUSED_METHODS:
            engine.__dict__.pop(n, None)
    _fused_engines.clear()
