"""Host side of the B200 backend: owns the native context and maps tensors onto the C ABI.

Mirrors the roles of the reference's NTTContext + MontgomeryContext + RnsPartition + constant pool
(tiberate/context/*.py) for the hot path, but every prime-dependent constant is derived inside
libtb200 (csrc/tb200.cu: tb200_ctx_create), lives in ordinary device memory (no 64-prime cap, not
process-global) and is addressed by explicit prime index.

Tensors are `torch.int64` CUDA tensors in the reference layout ([limbs, N], optionally with a
leading batch dimension).  NumPy arrays are accepted only so that tests/emu can drive the
host-compiled kernels; the product always runs on CUDA tensors.
"""

from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _native
from ._native import ExplicitConsts, Ksk, Poly, Tb200Error

R_BITS = 62
R = 1 << R_BITS


def _ptr(x) -> int:
    if x is None:
        return 0
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()


def _strides(x):
    if isinstance(x, np.ndarray):
        return [s // x.itemsize for s in x.strides]
    return list(x.stride())


def _is_int64(x) -> bool:
    if isinstance(x, np.ndarray):
        return x.dtype == np.int64
    import torch

    return x.dtype == torch.int64


def poly(x, N: int, batched: bool | None = None) -> Poly:
    """Describe a [rows, N] or [batch, rows, N] int64 tensor (row/batch strides may be views)."""
    if not _is_int64(x):
        raise Tb200Error(f"expected an int64 tensor, got {x.dtype}")
    st = _strides(x)
    nd = len(st)
    if x.shape[-1] != N or (x.shape[-1] > 1 and st[-1] != 1):
        raise Tb200Error(f"last dimension must be N={N} and contiguous, got shape {tuple(x.shape)} strides {st}")
    if nd == 1:
        return Poly(_ptr(x), 0, 0)
    if nd == 2:
        return Poly(_ptr(x), 0, st[0])
    if nd == 3:
        return Poly(_ptr(x), st[0], st[1])
    raise Tb200Error(f"expected 1, 2 or 3 dimensions, got {nd}")


def _stream(x) -> int:
    if isinstance(x, np.ndarray) or not x.is_cuda:
        return 0
    import torch

    return torch.cuda.current_stream(x.device).cuda_stream


class KeySwitchKeyView:
    """Pointers of a key-switch key: per global digit group (b, a), each [P, N]. Keeps the tensors alive."""

    def __init__(self, parts, N: int):
        self.parts = parts
        k = Ksk()
        k.num_groups = len(parts)
        rs = None
        for g, part in enumerate(parts):
            if part is None:
                continue
            b, a = part
            for t in (b, a):
                st = _strides(t)
                if len(st) != 2 or st[1] != 1 or t.shape[1] != N:
                    raise Tb200Error("key-switch key parts must be [P, N] row-contiguous int64 tensors")
                if rs is None:
                    rs = st[0]
                elif rs != st[0]:
                    raise Tb200Error("all key-switch key parts must share one row stride")
            k.b[g] = _ptr(b)
            k.a[g] = _ptr(a)
        k.row_stride = rs if rs is not None else N
        self.c = k
        self.rows = next((part[0].shape[0] for part in parts if part is not None), 0)
        for part in parts:
            if part is not None and (part[0].shape[0] != self.rows or part[1].shape[0] != self.rows):
                raise Tb200Error("all key-switch key parts must have the same number of rows")


class Tb200Context:
    """One per (parameter set, GPU).  `lib` defaults to the CUDA library (fails loudly if absent)."""

    def __init__(self, logN: int, q, num_special: int, scale_bits: int = 40, device: int = 0, lib=None,
                 rank: int = 0, world: int = 1):
        """`q` is always the GLOBAL prime chain.  With world > 1 the context is limb-sharded: it holds the
        ordinary primes of the digit groups `rank` owns plus the special primes (see include/tb200.h)."""
        self.lib = lib if lib is not None else _native.get_lib()
        self.rank, self.world = int(rank), int(world)
        self.logN, self.N = int(logN), 1 << int(logN)
        self.q = [int(x) for x in q]
        self.P, self.K = len(self.q), int(num_special)
        self.num_ordinary = self.P - self.K
        self.num_scales = self.num_ordinary - 1
        self.scale_bits = int(scale_bits)
        self.device = int(device)
        qa = np.ascontiguousarray(self.q, dtype=np.int64)
        if self.world == 1:
            h = self.lib.tb200_ctx_create(self.device, self.logN, self.P, self.K, qa.ctypes.data, self.scale_bits)
        else:
            h = self.lib.tb200_ctx_create_sharded(self.device, self.logN, self.P, self.K, qa.ctypes.data,
                                                  self.scale_bits, self.rank, self.world)
        if not h:
            raise Tb200Error("tb200_ctx_create failed: " + self.lib.tb200_last_error().decode(errors="replace"))
        self.h = C.c_void_p(h)
        info = (C.c_int32 * 8)()
        self.lib.check(self.lib.tb200_ctx_info(self.h, info), "ctx_info")
        self.LA, self.LB, self.num_groups0 = info[4], info[5], info[7]
        self.P_global, self.q_global = self.P, list(self.q)
        if self.world > 1:  # from here on P / q / num_ordinary describe the LOCAL rows
            self.P = info[2]
            ids = (C.c_int32 * self.P)()
            self.lib.check(self.lib.tb200_ctx_local_primes(self.h, ids), "local_primes")
            self.local_prime_ids = list(ids)
            self.q = [self.q_global[i] for i in self.local_prime_ids]
            self.num_ordinary = self.P - self.K
        else:
            self.local_prime_ids = list(range(self.P))
        # host mirror of the Montgomery constants (mont_context.py:26-57) for callers that need them
        self.k = [(R * pow(R, -1, qi) - 1) // qi for qi in self.q]
        self.Rs = [R * R % qi for qi in self.q]

    def close(self):
        if getattr(self, "h", None):
            self.lib.tb200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection (tests) -----------------------------------------------------------
    def prime_consts(self) -> np.ndarray:
        out = np.zeros((self.P, 8), dtype=np.int64)
        self.lib.check(self.lib.tb200_ctx_get_prime_consts(self.h, out.ctypes.data), "get_prime_consts")
        return out

    def twiddles(self, inverse: bool, prime: int) -> np.ndarray:
        out = np.zeros(self.N, dtype=np.int64)
        self.lib.check(self.lib.tb200_ctx_get_twiddles(self.h, int(inverse), prime, out.ctypes.data), "get_twiddles")
        return out

    def set_chunk(self, chunk: int):
        self.lib.check(self.lib.tb200_ctx_set_chunk(self.h, int(chunk)), "set_chunk")

    def set_fast(self, on: bool):
        """Internal transforms of the fused engine calls: mod-q path (default) or exact op kernels."""
        self.lib.check(self.lib.tb200_ctx_set_fast(self.h, int(bool(on))), "set_fast")

    def set_f64_share(self, eighths: int):
        """Share (0..8 eighths) of the small-prime limbs transformed on the FP64 pipe (mod-q path)."""
        self.lib.check(self.lib.tb200_ctx_set_f64_share(self.h, int(eighths)), "set_f64_share")

    TUNE_FUSED_CORE, TUNE_SIDE_ROWS, TUNE_FUSED_MODDOWN, TUNE_STREAM_WS, TUNE_SUM_NTT, TUNE_FUSED_TENSOR = 0, 1, 2, 3, 4, 5

    def set_tuning(self, knob: int, value: int):
        """Scheduling knobs of the mod-q path (include/tb200.h: enum tb200_tuning); results never change."""
        self.lib.check(self.lib.tb200_ctx_set_tuning(self.h, int(knob), int(value)), "set_tuning")

    # ---- level helpers -------------------------------------------------------------------
    def rows_at(self, level: int, with_special: bool = False) -> int:
        return (self.P if with_special else self.num_ordinary) - level

    def prime0_for(self, rows: int, sp_prime_len: int) -> int:
        """a tensor of `rows` rows called with sp_prime_len addresses the
        constant pool right-aligned: row i -> prime P - rows - sp_prime_len + i."""
        p0 = self.P - rows - sp_prime_len
        if p0 < self.P - 64:  # the reference would read another constant's region of its pool
            raise Tb200Error(f"{rows} rows with sp_prime_len={sp_prime_len} exceed the reference's 64-slot pool")
        # p0 < 0: the leading rows use the zero padding of the pool (pointwise ops only; the C side checks)
        return p0

    # ---- op layer ------------------------------------------------------------------------
    def pointwise(self, op: int, a, b=None, out=None, prime0: int = 0, scal=None, ec: ExplicitConsts | None = None):
        N = self.N
        pa = poly(a, N)
        out = a if out is None else out
        po = poly(out, N)
        rows = out.shape[-2] if len(out.shape) >= 2 else 1
        batch = out.shape[0] if len(out.shape) == 3 else 1
        pb = poly(b, N) if b is not None else None
        rc = self.lib.tb200_pointwise(self.h, op, rows, batch, prime0, C.byref(pa),
                                      C.byref(pb) if pb is not None else None, _ptr(scal),
                                      C.byref(ec) if ec is not None else None, C.byref(po), _stream(out))
        self.lib.check(rc, f"pointwise op {op}")
        return out

    def add_many(self, stacked, out, prime0: int, pairwise: bool):
        K, rows, N = stacked.shape
        assert N == self.N
        rc = self.lib.tb200_add_many(self.h, int(pairwise), K, rows, prime0, _ptr(stacked), _ptr(out), _stream(out))
        self.lib.check(rc, "add_many")
        return out

    def ntt(self, a, prime0: int, enter: bool, rows: int | None = None):
        pa = poly(a, self.N)
        batch = a.shape[0] if len(a.shape) == 3 else 1
        rows = a.shape[-2] if rows is None else rows
        self.lib.check(self.lib.tb200_ntt(self.h, rows, batch, prime0, C.byref(pa), int(enter), _stream(a)), "ntt")
        return a

    def intt(self, a, prime0: int, mode: int, rows: int | None = None):
        pa = poly(a, self.N)
        batch = a.shape[0] if len(a.shape) == 3 else 1
        rows = a.shape[-2] if rows is None else rows
        self.lib.check(self.lib.tb200_intt(self.h, rows, batch, prime0, C.byref(pa), int(mode), _stream(a)), "intt")
        return a

    def rescale_rows(self, a, prime0: int, scales, rescaler, round_at: int, exact: bool):
        pa = poly(a, self.N)
        rc = self.lib.tb200_rescale_rows(self.h, a.shape[0], prime0, C.byref(pa), _ptr(scales), _ptr(rescaler),
                                         int(round_at), int(exact), _stream(a))
        self.lib.check(rc, "rescale_rows")
        return a

    def extend(self, rns_len: int, prime0: int, state, l_enter, l_enter_offset: int, out):
        alpha = state.shape[0]
        le_stride = _strides(l_enter)[0] if (l_enter is not None and alpha > 1) else 0
        rc = self.lib.tb200_extend(self.h, rns_len, prime0, alpha, _ptr(state), _strides(state)[0],
                                   _ptr(l_enter) if alpha > 1 else 0, le_stride, int(l_enter_offset), _ptr(out),
                                   _strides(out)[0], _stream(out))
        self.lib.check(rc, "extend")
        return out

    def codec_rotate(self, a, perm, two_q, out):
        rc = self.lib.tb200_codec_rotate(self.h, a.shape[0], C.byref(poly(a, self.N)), _ptr(perm), _ptr(two_q),
                                         C.byref(poly(out, self.N)), _stream(out))
        self.lib.check(rc, "codec_rotate")
        return out

    def divide_by_p(self, level: int, c, p, out):
        rc = self.lib.tb200_divide_by_p(self.h, level, C.byref(poly(c, self.N)), C.byref(poly(p, self.N)),
                                        C.byref(poly(out, self.N)), _stream(out))
        self.lib.check(rc, "divide_by_p")
        return out

    # ---- packed wire format (include/tb200.h) --------------------------------------------------
    def narrow_rows(self, prime0: int, rows: int) -> int:
        """how many of the `rows` limbs starting at prime index prime0 are scale-prime limbs (below 2^41; a
        prefix: the chain is [scale primes..., base, special...])"""
        n = 0
        while n < rows and self.q[prime0 + n] < (1 << 41):
            n += 1
        return n

    @property
    def packed_row_bytes(self) -> int:
        """bytes of one packed limb: 5 N + N / 8 (include/tb200.h, packed wire format)"""
        return 5 * self.N + self.N // 8

    def unpack41(self, packed, dst, prime0: int):
        """packed: uint8 [(B,) rows, packed_row_bytes] (last dim contiguous) -> dst int64 [(B,) rows, N]"""
        rows, batch = dst.shape[-2], self._batch(dst)
        if packed.shape[-1] != self.packed_row_bytes or packed.shape[-2] != rows or _strides(packed)[-1] != 1:
            raise Tb200Error(f"unpack41: packed must be [.., {rows}, {self.packed_row_bytes}] bytes, got {tuple(packed.shape)}")
        st = _strides(packed)
        rc = self.lib.tb200_unpack41(self.h, rows, batch, prime0, _ptr(packed), st[0] if len(st) == 3 else 0, st[-2],
                                     self._pp(dst), _stream(dst))
        self.lib.check(rc, "unpack41")
        return dst

    def pack41(self, src, packed, prime0: int):
        """src int64 [(B,) rows, N] canonical -> packed uint8 [(B,) rows, packed_row_bytes]"""
        rows, batch = src.shape[-2], self._batch(src)
        if packed.shape[-1] != self.packed_row_bytes or packed.shape[-2] != rows or _strides(packed)[-1] != 1:
            raise Tb200Error(f"pack41: packed must be [.., {rows}, {self.packed_row_bytes}] bytes, got {tuple(packed.shape)}")
        st = _strides(packed)
        rc = self.lib.tb200_pack41(self.h, rows, batch, prime0, self._pp(src), _ptr(packed), st[0] if len(st) == 3 else 0,
                                   st[-2], _stream(src))
        self.lib.check(rc, "pack41")
        return packed

    # ---- engine layer --------------------------------------------------------------------
    @staticmethod
    def _batch(x) -> int:
        return x.shape[0] if len(x.shape) == 3 else 1

    def _pp(self, x):
        return C.byref(poly(x, self.N)) if x is not None else None

    def _rows(self, level: int) -> int:
        """limb rows of a polynomial at `level` in this context (local rows when limb-sharded)"""
        if self.world == 1:
            return self.num_ordinary - level
        return len(self.local_rows(level))

    def _shapes(self, what: str, level: int, *specs, key: KeySwitchKeyView | None = None):
        """The C entry points take raw pointers and derive every extent from `level`: check here that each
        operand really has that many limb rows, and that all of them agree on batch size and device.
        specs: (name, tensor or None, expected rows, may_broadcast_batch)."""
        if not 0 <= level < self.P_global - self.K:
            raise Tb200Error(f"{what}: level {level} outside [0, {self.P_global - self.K})")
        batch = dev = None
        for name, x, rows, *opt in specs:
            if x is None:
                continue
            nd = len(x.shape)
            if nd not in (2, 3):
                raise Tb200Error(f"{what}: {name} must be [limbs, N] or [batch, limbs, N], got shape {tuple(x.shape)}")
            if x.shape[-2] != rows:
                raise Tb200Error(f"{what}: {name} has {x.shape[-2]} limb rows, level {level} needs {rows} "
                                 f"(shape {tuple(x.shape)}; ciphertexts that include the special limbs are not "
                                 f"accepted here)")
            b = x.shape[0] if nd == 3 else 1
            if not (opt and opt[0] and nd == 2):
                if batch is None:
                    batch = b
                elif b != batch:
                    raise Tb200Error(f"{what}: {name} has batch {b}, the other operands {batch}")
            d = None if isinstance(x, np.ndarray) else x.device
            if d is not None and not (d.type == "cpu" and "emu" in getattr(self.lib, "path", "")):  # tests/emu: host tensors
                if d.type != "cuda" or (d.index if d.index is not None else 0) != self.device:
                    raise Tb200Error(f"{what}: {name} lives on {d}, the context on cuda:{self.device}")
                dev = d
        if key is not None and key.rows != len(self.local_prime_ids):
            raise Tb200Error(f"{what}: key-switch key parts have {key.rows} rows, the context has "
                             f"{len(self.local_prime_ids)} primes")
        _ = dev

    def rescale(self, level: int, in0, in1, out0, out1, exact: bool = True):
        r = self._rows(level)
        self._shapes("rescale", level, ("in0", in0, r), ("in1", in1, r), ("out0", out0, r - 1), ("out1", out1, r - 1))
        rc = self.lib.tb200_rescale(self.h, level, self._batch(in0), self._pp(in0), self._pp(in1), self._pp(out0),
                                    self._pp(out1), int(exact), _stream(out0))
        self.lib.check(rc, "rescale")

    def keyswitch(self, level: int, a, ksk: KeySwitchKeyView, out0, out1):
        r = self._rows(level)
        self._shapes("keyswitch", level, ("a", a, r), ("out0", out0, r), ("out1", out1, r), key=ksk)
        rc = self.lib.tb200_keyswitch(self.h, level, self._batch(a), self._pp(a), C.byref(ksk.c), self._pp(out0),
                                      self._pp(out1), _stream(out0))
        self.lib.check(rc, "keyswitch")

    # ---- key switch in two halves (the seam for limb sharding) ---------------------------------
    def ks_state_info(self, level: int):
        """(state_rows, segment_row0, segment_rows, local_ordinary_rows) at `level`."""
        out = (C.c_int32 * 4)()
        self.lib.check(self.lib.tb200_ks_state_info(self.h, level, out), "ks_state_info")
        return tuple(out)

    def local_rows(self, level: int, with_special: bool = False):
        """Global prime ids of the local rows alive at `level`."""
        ids = [g for g in self.local_prime_ids[: self.num_ordinary] if g >= level]
        return ids + (self.local_prime_ids[self.num_ordinary:] if with_special else [])

    def ks_digits(self, level: int, a, state):
        rc = self.lib.tb200_ks_digits(self.h, level, self._batch(state), self._pp(a), self._pp(state), _stream(state))
        self.lib.check(rc, "ks_digits")

    def ks_finish(self, level: int, state, ksk: KeySwitchKeyView, out0, out1, add0=None, add1=None, tail: int = 0):
        r = self._rows(level)
        self._shapes("ks_finish", level, ("state", state, self.ks_state_info(level)[0]), ("out0", out0, r),
                     ("out1", out1, r), ("add0", add0, r), ("add1", add1, r), key=ksk)
        rc = self.lib.tb200_ks_finish(self.h, level, self._batch(state), self._pp(state), C.byref(ksk.c),
                                      self._pp(add0), self._pp(add1), self._pp(out0), self._pp(out1), int(tail),
                                      _stream(out0))
        self.lib.check(rc, "ks_finish")

    def ks_modup(self, level: int, state, which: int = 0):
        """ModUp (extend + forward pass A) of the digit groups selected by `which`: 0 all, 1 the groups this
        rank owns, 2 the others; into the context workspace (followed by ks_core)."""
        self._shapes("ks_modup", level, ("state", state, self.ks_state_info(level)[0]))
        rc = self.lib.tb200_ks_modup(self.h, level, self._batch(state), self._pp(state), int(which), _stream(state))
        self.lib.check(rc, "ks_modup")

    def ks_core(self, level: int, state, ksk: KeySwitchKeyView, out0, out1, add0=None, add1=None, tail: int = 0):
        r = self._rows(level)
        self._shapes("ks_core", level, ("state", state, self.ks_state_info(level)[0]), ("out0", out0, r),
                     ("out1", out1, r), ("add0", add0, r), ("add1", add1, r), key=ksk)
        rc = self.lib.tb200_ks_core(self.h, level, self._batch(state), self._pp(state), C.byref(ksk.c),
                                    self._pp(add0), self._pp(add1), self._pp(out0), self._pp(out1), int(tail),
                                    _stream(out0))
        self.lib.check(rc, "ks_core")

    # ---- the same with the special limbs sharded too (include/tb200.h: tb200_ks_core_sp) -------
    def ks_sp_info(self):
        """(rows of the sp buffer, rows per rank segment, s0, s1): this rank computes special limbs [s0, s1)."""
        out = (C.c_int32 * 4)()
        self.lib.check(self.lib.tb200_ks_sp_info(self.h, out), "ks_sp_info")
        return tuple(out)

    def _sp_check(self, what, sp, batch):
        rows = self.ks_sp_info()[0]
        want = [batch * 2 * self.N, 2 * self.N, self.N, 1]
        if tuple(sp.shape) != (rows, batch, 2, self.N) or not _is_int64(sp) or _strides(sp) != want:
            raise Tb200Error(f"{what}: sp must be a contiguous int64 tensor [{rows}, {batch}, 2, {self.N}]")

    def ks_core_sp(self, level: int, batch: int, ksk: KeySwitchKeyView, sp):
        """Key sums of this rank's share of the special limbs (after ks_modup(which + 4)) -> its segment of sp."""
        self._sp_check("ks_core_sp", sp, batch)
        rc = self.lib.tb200_ks_core_sp(self.h, level, int(batch), C.byref(ksk.c), C.c_void_p(_ptr(sp)), _stream(sp))
        self.lib.check(rc, "ks_core_sp")

    def ks_core_ord(self, level: int, batch: int, ksk: KeySwitchKeyView, like):
        """Key sums of the local ordinary limbs (they stay in the context workspace until ks_moddown)."""
        rc = self.lib.tb200_ks_core_ord(self.h, level, int(batch), C.byref(ksk.c), _stream(like))
        self.lib.check(rc, "ks_core_ord")

    def ks_moddown(self, level: int, sp, out0, out1, add0=None, add1=None, tail: int = 0):
        r = self._rows(level)
        self._shapes("ks_moddown", level, ("out0", out0, r), ("out1", out1, r), ("add0", add0, r), ("add1", add1, r))
        batch = self._batch(out0)
        self._sp_check("ks_moddown", sp, batch)
        rc = self.lib.tb200_ks_moddown(self.h, level, batch, C.c_void_p(_ptr(sp)), self._pp(add0), self._pp(add1),
                                       self._pp(out0), self._pp(out1), int(tail), _stream(out0))
        self.lib.check(rc, "ks_moddown")

    def cc_mult_relin(self, level: int, a0, a1, b0, b1, evk: KeySwitchKeyView, out0, out1, pre_rescale: bool = True):
        r = self._rows(level)
        ro = r - (1 if pre_rescale else 0)
        self._shapes("cc_mult_relin", level, ("a0", a0, r), ("a1", a1, r), ("b0", b0, r), ("b1", b1, r),
                     ("out0", out0, ro), ("out1", out1, ro), key=evk)
        rc = self.lib.tb200_cc_mult_relin(self.h, level, self._batch(a0), self._pp(a0), self._pp(a1), self._pp(b0),
                                          self._pp(b1), C.byref(evk.c), self._pp(out0), self._pp(out1),
                                          int(pre_rescale), _stream(out0))
        self.lib.check(rc, "cc_mult_relin")

    def cc_mult_triplet(self, level: int, a0, a1, b0, b1, d0, d1, d2, pre_rescale: bool = True):
        r = self._rows(level)
        ro = r - (1 if pre_rescale else 0)
        self._shapes("cc_mult_triplet", level, ("a0", a0, r), ("a1", a1, r), ("b0", b0, r), ("b1", b1, r),
                     ("d0", d0, ro), ("d1", d1, ro), ("d2", d2, ro))
        rc = self.lib.tb200_cc_mult_triplet(self.h, level, self._batch(a0), self._pp(a0), self._pp(a1), self._pp(b0),
                                            self._pp(b1), self._pp(d0), self._pp(d1), self._pp(d2), int(pre_rescale),
                                            _stream(d0))
        self.lib.check(rc, "cc_mult_triplet")

    def relinearize(self, level: int, d0, d1, d2, evk: KeySwitchKeyView, out0, out1):
        r = self._rows(level)
        self._shapes("relinearize", level, ("d0", d0, r), ("d1", d1, r), ("d2", d2, r), ("out0", out0, r),
                     ("out1", out1, r), key=evk)
        rc = self.lib.tb200_relinearize(self.h, level, self._batch(d0), self._pp(d0), self._pp(d1), self._pp(d2),
                                        C.byref(evk.c), self._pp(out0), self._pp(out1), _stream(out0))
        self.lib.check(rc, "relinearize")

    def rotate(self, level: int, galois: int, c0, c1, rotk: KeySwitchKeyView | None, out0, out1):
        r = self._rows(level)
        self._shapes("rotate", level, ("c0", c0, r), ("c1", c1, r), ("out0", out0, r), ("out1", out1, r), key=rotk)
        rc = self.lib.tb200_rotate(self.h, level, self._batch(c0), int(galois), self._pp(c0), self._pp(c1),
                                   C.byref(rotk.c) if rotk is not None else None, self._pp(out0), self._pp(out1),
                                   _stream(out0))
        self.lib.check(rc, "rotate")

    def rotate_hoisted(self, level: int, galois, c0, c1, rotks, out0, out1):
        """Rotations galois[r] (keys rotks[r]) of the same ciphertext(s) with the ModUp shared (extension beyond
        the reference, see include/tb200.h).  out0 / out1: [R, (B,) L, N] dense over the rotation axis."""
        R = len(galois)
        if len(rotks) != R or out0.shape[0] != R or out1.shape[0] != R:
            raise Tb200Error("rotate_hoisted: one key and one output slice per rotation")
        r = self._rows(level)
        self._shapes("rotate_hoisted", level, ("c0", c0, r), ("c1", c1, r), ("out0", out0[0], r), ("out1", out1[0], r),
                     key=rotks[0])
        for k in rotks[1:]:
            self._shapes("rotate_hoisted", level, key=k)
        if _strides(out0)[0] != _strides(out1)[0]:
            raise Tb200Error("rotate_hoisted: out0 and out1 need the same stride over the rotation axis")
        g = (C.c_int64 * R)(*[int(x) for x in galois])
        kp = (C.c_void_p * R)(*[C.cast(C.pointer(k.c), C.c_void_p) for k in rotks])
        rc = self.lib.tb200_rotate_hoisted(self.h, level, self._batch(c0), R, g, self._pp(c0), self._pp(c1), kp,
                                           self._pp(out0[0]), self._pp(out1[0]), _strides(out0)[0], _stream(out0))
        self.lib.check(rc, "rotate_hoisted")

    def switch_key(self, level: int, c0, c1, ksk: KeySwitchKeyView, out0, out1):
        r = self._rows(level)
        self._shapes("switch_key", level, ("c0", c0, r), ("c1", c1, r), ("out0", out0, r), ("out1", out1, r), key=ksk)
        rc = self.lib.tb200_switch_key(self.h, level, self._batch(c0), self._pp(c0), self._pp(c1), C.byref(ksk.c),
                                       self._pp(out0), self._pp(out1), _stream(out0))
        self.lib.check(rc, "switch_key")

    def pc_mult(self, level: int, pt, c0, c1, out0, out1, post_rescale: bool = True):
        r = self._rows(level)
        ro = r - (1 if post_rescale else 0)
        self._shapes("pc_mult", level, ("pt", pt, r, True), ("c0", c0, r), ("c1", c1, r), ("out0", out0, ro),
                     ("out1", out1, ro))
        ppt = poly(pt, self.N)
        if len(pt.shape) == 2:
            ppt.batch_stride = 0
        rc = self.lib.tb200_pc_mult(self.h, level, self._batch(c0), C.byref(ppt), self._pp(c0), self._pp(c1),
                                    self._pp(out0), self._pp(out1), int(post_rescale), _stream(out0))
        self.lib.check(rc, "pc_mult")

    def cc_addsub(self, level: int, sub: bool, a0, a1, b0, b1, out0, out1):
        r = self._rows(level)
        self._shapes("cc_addsub", level, ("a0", a0, r), ("a1", a1, r), ("b0", b0, r), ("b1", b1, r),
                     ("out0", out0, r), ("out1", out1, r))
        rc = self.lib.tb200_cc_addsub(self.h, level, self._batch(a0), int(sub), self._pp(a0), self._pp(a1),
                                      self._pp(b0), self._pp(b1), self._pp(out0), self._pp(out1), _stream(out0))
        self.lib.check(rc, "cc_addsub")


def galois_element(N: int, delta: int) -> int:
    """ckks_engine.py:1817-1821: leap = (3^delta - 1)/2 mod 2N, p = 2*leap + 1 = 3^delta mod 2N."""
    d = delta % N
    leap = (3 ** d - 1) // 2 % (2 * N)
    return (2 * leap + 1) % (2 * N)


_ = math
