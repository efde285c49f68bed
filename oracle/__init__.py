"""CPU oracle for the CKKS RNS hot path of tiberate-fhe.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only as the checker or
the reported CPU baseline.  The product (``tiberate_fhe_b200``) never imports it.

Parity pins (the reference has no golden vectors and no CPU path):
  * big-int identities (tests/test_oracle_*.py);
  * constants produced by the reference's own Python context code, imported in the build
    container (tests/golden/make_ctx_golden.py -> tests/golden/ctx_*.json);
  * op outputs of the reference's own CUDA extension run on a B200
    (tests/golden/make_ref_golden.py -> tests/golden/ref_*.npz).
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libckks_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/ckks_oracle.c (gcc, seconds)."""
    src = os.path.join(_HERE, "ckks_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64 = ctypes.c_int64
        for name in ("orc_mm_halves", "orc_mm_closed"):
            f = getattr(_lib, name)
            f.restype = i64
            f.argtypes = [i64, i64, i64, i64]
        for name in ("orc_mr_halves", "orc_mr_closed"):
            f = getattr(_lib, name)
            f.restype = i64
            f.argtypes = [i64, i64, i64]
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return ctypes.c_void_p(a.ctypes.data)


def _c(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.int64)


_L = ctypes.c_long
_I = ctypes.c_int
_Q = ctypes.c_int64


def set_threads(n: int) -> None:
    """OpenMP threads used by the C restatement (the default is libgomp's: OMP_NUM_THREADS or all cores)."""
    lib().orc_set_threads(int(n))


def call(name: str, *args):
    """Call a void C function; numpy arrays become pointers, ints become long."""
    conv = []
    for a in args:
        if isinstance(a, np.ndarray):
            conv.append(_p(a))
        elif isinstance(a, (_L, _I, _Q, ctypes.c_size_t)):
            conv.append(a)
        else:
            conv.append(_L(int(a)))
    f = getattr(lib(), name)
    f.restype = None
    f(*conv)
