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


def install_as_tiberate_backend(engine) -> list[Tb200Context]:
    """`engine` is a constructed reference CkksEngine; one Tb200Context per engine device is created
    from that device's local prime list (rnsPart.d_special) and made current."""
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
    return ctxs


def uninstall() -> None:
    import tiberate.libs.wrapper as ref_wrapper

    for (mod_name, n), fn in _saved.items():
        setattr(importlib.import_module(f"tiberate.libs.wrapper.{mod_name}"), n, fn)
    _saved.clear()
