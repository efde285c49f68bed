"""Drop-in mirrors of the reference's generated op wrappers (tiberate/libs/wrapper/{mont_ops,
ntt2_ops,he_ops,csprng_ops,const_pool}.py): same function names, argument order and meaning, same in-place / functional
behaviour, every tensor argument a per-device list.  They forward to libtb200 through the C ABI.

The reference keeps its per-prime constants in a process-global __constant__ pool per device
(csrc/ops/cuda/constant_mem.cuh); here the equivalent is the *current context* of a device:
`set_context(ctx)` (done by Tb200-backed engines and by backend.install_as_tiberate_backend).
"""

from __future__ import annotations

from ..context import Tb200Context, Tb200Error

import threading

_current: dict[int, Tb200Context] = {}
_override = threading.local()  # set while an engine-bound wrapper function runs (bind)


def set_context(ctx: Tb200Context) -> None:
    _current[ctx.device] = ctx


def clear_contexts() -> None:
    _current.clear()


class _Bound:
    """A wrapper module whose functions run on one given context (an engine's own), whatever the
    device-current context is: two engines with different prime chains can share a GPU."""

    def __init__(self, module, ctx: Tb200Context):
        self._module, self._ctx = module, ctx

    def __getattr__(self, name):
        fn = getattr(self._module, name)
        if not callable(fn):
            return fn
        ctx = self._ctx

        def call(*args, **kwargs):
            prev = getattr(_override, "ctx", None)
            _override.ctx = ctx
            try:
                return fn(*args, **kwargs)
            finally:
                _override.ctx = prev

        call.__name__ = name
        setattr(self, name, call)
        return call


def bind(module, ctx: Tb200Context) -> _Bound:
    return _Bound(module, ctx)


def context_for(t) -> Tb200Context:
    ctx = getattr(_override, "ctx", None)
    if ctx is not None:
        return ctx
    idx = t.device.index if t.device.index is not None else 0
    try:
        return _current[idx]
    except KeyError:
        raise Tb200Error(
            f"no tb200 context is current on cuda:{idx}; construct an engine or call wrapper.set_context(ctx)"
        ) from None


from . import const_pool, csprng_ops, he_ops, mont_ops, ntt2_ops  # noqa: E402,F401
