"""Drop-in mirrors of the reference's generated op wrappers (tiberate/libs/wrapper/{mont_ops,
ntt2_ops,he_ops,csprng_ops,const_pool}.py): same function names, argument order and meaning, same in-place / functional
behaviour, every tensor argument a per-device list.  They forward to libtb200 through the C ABI.

The reference keeps its per-prime constants in a process-global __constant__ pool per device
(csrc/ops/cuda/constant_mem.cuh); here the equivalent is the *current context* of a device:
`set_context(ctx)` (done by Tb200-backed engines and by backend.install_as_tiberate_backend).
"""

from __future__ import annotations

from ..context import Tb200Context, Tb200Error

_current: dict[int, Tb200Context] = {}


def set_context(ctx: Tb200Context) -> None:
    _current[ctx.device] = ctx


def clear_contexts() -> None:
    _current.clear()


def context_for(t) -> Tb200Context:
    idx = t.device.index if t.device.index is not None else 0
    try:
        return _current[idx]
    except KeyError:
        raise Tb200Error(
            f"no tb200 context is current on cuda:{idx}; construct an engine or call wrapper.set_context(ctx)"
        ) from None


from . import const_pool, csprng_ops, he_ops, mont_ops, ntt2_ops  # noqa: E402,F401
