"""tiberate_fhe_b200 -- B200 (sm_100a) backend for the CKKS RNS polynomial hot path behind
tiberate-fhe's CkksEngine (NTT/iNTT, Montgomery pointwise, rescale, key-switch, automorphism).

Host code is Python (the reference's host language); the compute path is libtb200.so
(hand-written CUDA, C ABI in include/tb200.h).  No CPU fallback.
"""

from ._native import Tb200Error, get_lib  # noqa: F401
from .context import KeySwitchKeyView, Tb200Context, galois_element  # noqa: F401



def __getattr__(name):  # torch is imported only when the engine layer is used
    if name == "CkksEngine":
        from .ckks_engine import CkksEngine

        return CkksEngine
    if name == "Csprng":
        from .rng import Csprng

        return Csprng
    raise AttributeError(name)


__version__ = "0.1.0"
