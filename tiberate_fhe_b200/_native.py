"""ctypes binding of libtb200.so (the C ABI in include/tb200.h).

The product path has no CPU fallback: `get_lib()` raises if the CUDA library is missing or no
CUDA device is usable.  (tests/emu builds the same sources for the host to check kernel indexing
without a GPU; only the tests bind that build, through `Lib(path)` explicitly.)
"""

from __future__ import annotations

import ctypes as C
import os

MAX_GROUPS = 32

_HERE = os.path.dirname(os.path.abspath(__file__))
# TB200_LIB: developer override used to A/B kernel build variants on the GPU box
LIB_PATH = os.environ.get("TB200_LIB") or os.path.join(_HERE, "lib", "libtb200.so")


class Poly(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("batch_stride", C.c_int64), ("row_stride", C.c_int64)]


class Ksk(C.Structure):
    _fields_ = [
        ("num_groups", C.c_int32),
        ("reserved", C.c_int32),
        ("row_stride", C.c_int64),
        ("b", C.c_void_p * MAX_GROUPS),
        ("a", C.c_void_p * MAX_GROUPS),
    ]


class ExplicitConsts(C.Structure):
    _fields_ = [("ql", C.c_void_p), ("qh", C.c_void_p), ("kl", C.c_void_p), ("kh", C.c_void_p),
                ("two_q", C.c_void_p)]


PP = C.POINTER(Poly)
_i, _i64, _vp = C.c_int, C.c_int64, C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/tb200.h
SIGNATURES = {
    "tb200_ctx_create": (_vp, [_i, _i, _i, _i, _vp, _i]),
    "tb200_ctx_create_sharded": (_vp, [_i, _i, _i, _i, _vp, _i, _i, _i]),
    "tb200_ctx_local_primes": (_i, [_vp, _vp]),
    "tb200_ctx_destroy": (None, [_vp]),
    "tb200_last_error": (C.c_char_p, []),
    "tb200_version": (C.c_char_p, []),
    "tb200_ctx_get_prime_consts": (_i, [_vp, _vp]),
    "tb200_ctx_get_twiddles": (_i, [_vp, _i, _i, _vp]),
    "tb200_ctx_info": (_i, [_vp, _vp]),
    "tb200_ctx_set_chunk": (_i, [_vp, _i]),
    "tb200_ctx_set_fast": (_i, [_vp, _i]),
    "tb200_ctx_set_f64_share": (_i, [_vp, _i]),
    "tb200_ctx_set_tuning": (_i, [_vp, _i, _i]),
    "tb200_pointwise": (_i, [_vp, _i, _i, _i, _i, PP, PP, _vp, C.POINTER(ExplicitConsts), PP, _vp]),
    "tb200_add_many": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "tb200_ntt": (_i, [_vp, _i, _i, _i, PP, _i, _vp]),
    "tb200_intt": (_i, [_vp, _i, _i, _i, PP, _i, _vp]),
    "tb200_rescale_rows": (_i, [_vp, _i, _i, PP, _vp, _vp, _i64, _i, _vp]),
    "tb200_extend": (_i, [_vp, _i, _i, _i, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _vp]),
    "tb200_codec_rotate": (_i, [_vp, _i, PP, _vp, _vp, PP, _vp]),
    "tb200_divide_by_p": (_i, [_vp, _i, PP, PP, PP, _vp]),
    "tb200_rescale": (_i, [_vp, _i, _i, PP, PP, PP, PP, _i, _vp]),
    "tb200_keyswitch": (_i, [_vp, _i, _i, PP, C.POINTER(Ksk), PP, PP, _vp]),
    "tb200_ks_state_info": (_i, [_vp, _i, _vp]),
    "tb200_ks_digits": (_i, [_vp, _i, _i, PP, PP, _vp]),
    "tb200_ks_finish": (_i, [_vp, _i, _i, PP, C.POINTER(Ksk), PP, PP, PP, PP, _i, _vp]),
    "tb200_ks_modup": (_i, [_vp, _i, _i, PP, _i, _vp]),
    "tb200_unpack41": (_i, [_vp, _i, _i, _i, _vp, C.c_int64, C.c_int64, PP, _vp]),
    "tb200_pack41": (_i, [_vp, _i, _i, _i, PP, _vp, C.c_int64, C.c_int64, _vp]),
    "tb200_ks_core": (_i, [_vp, _i, _i, PP, C.POINTER(Ksk), PP, PP, PP, PP, _i, _vp]),
    "tb200_ks_sp_info": (_i, [_vp, _vp]),
    "tb200_ks_core_sp": (_i, [_vp, _i, _i, C.POINTER(Ksk), _vp, _vp]),
    "tb200_ks_core_ord": (_i, [_vp, _i, _i, C.POINTER(Ksk), _vp]),
    "tb200_ks_moddown": (_i, [_vp, _i, _i, _vp, PP, PP, PP, PP, _i, _vp]),
    "tb200_cc_mult_relin": (_i, [_vp, _i, _i, PP, PP, PP, PP, C.POINTER(Ksk), PP, PP, _i, _vp]),
    "tb200_cc_mult_triplet": (_i, [_vp, _i, _i, PP, PP, PP, PP, PP, PP, PP, _i, _vp]),
    "tb200_relinearize": (_i, [_vp, _i, _i, PP, PP, PP, C.POINTER(Ksk), PP, PP, _vp]),
    "tb200_rotate": (_i, [_vp, _i, _i, _i64, PP, PP, C.POINTER(Ksk), PP, PP, _vp]),
    "tb200_rotate_hoisted": (_i, [_vp, _i, _i, _i, _vp, PP, PP, _vp, PP, PP, C.c_int64, _vp]),
    "tb200_switch_key": (_i, [_vp, _i, _i, PP, PP, C.POINTER(Ksk), PP, PP, _vp]),
    "tb200_pc_mult": (_i, [_vp, _i, _i, PP, PP, PP, PP, PP, _i, _vp]),
    "tb200_cc_addsub": (_i, [_vp, _i, _i, _i, PP, PP, PP, PP, PP, PP, _vp]),
    "tb200_chacha20": (_i, [_i, C.c_void_p, _i64, C.c_void_p, _i64, C.c_void_p]),
    "tb200_randint_fast": (_i, [_i, C.c_void_p, _i, _i64, C.c_void_p, _i64, _i64, C.c_void_p, C.c_void_p]),
    "tb200_discrete_gaussian_fast": (_i, [_i, C.c_void_p, _i64, C.c_void_p, _i, _i, _i64, C.c_void_p, C.c_void_p]),
    "tb200_randint": (_i, [_i, C.c_void_p, _i, _i64, C.c_void_p, C.c_void_p]),
    "tb200_discrete_gaussian": (_i, [_i, C.c_void_p, _i64, C.c_void_p, _i, _i, C.c_void_p]),
    "tb200_randround": (_i, [_i, C.c_void_p, C.c_void_p, _i64, C.c_void_p]),
    "tb200_launch_count": (_i64, []),
    "tb200_prof_enable": (None, [_i]),
    "tb200_prof_collect": (_i, [C.c_char_p, _i]),
}


class Tb200Error(RuntimeError):
    """Raised for every non-zero return code of the C ABI (the reference raises RuntimeError
    through TORCH_CHECK)."""


class Lib:
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise Tb200Error(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback."
            )
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(self.dll, name)  # AttributeError if the .so lacks a declared symbol
            f.restype = res
            f.argtypes = args

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.dll.tb200_last_error().decode(errors="replace")
            raise Tb200Error(f"{what} failed (code {rc}): {msg}")

    def __getattr__(self, name):
        return getattr(self.dll, name)


_lib = None


def get_lib() -> Lib:
    """The CUDA library; fails loudly when it is absent."""
    global _lib
    if _lib is None:
        _lib = Lib(LIB_PATH)
    return _lib
