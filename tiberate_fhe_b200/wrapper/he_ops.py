"""tiberate/libs/wrapper/he_ops.py mirror (schemas: csrc/ops/he_fused.cpp:80-99)."""

from __future__ import annotations

import torch

from . import context_for
from .mont_ops import PC_ADD


def pc_add_fused(ct, pt, sp_prime_len):
    outs = []
    for ci, pi in zip(ct, pt):
        ctx = context_for(ci)
        out = torch.empty_like(ci, memory_format=torch.contiguous_format)
        ctx.pointwise(PC_ADD, ci, pi, out, ctx.prime0_for(ci.size(0), sp_prime_len))
        outs.append(out)
    return outs


def _rescale(a, scales, rescaler, round_at, sp_prime_len, exact):
    for ai, si, ri in zip(a, scales, rescaler):
        if isinstance(ai, list) or ai is None or ai.numel() == 0:
            continue
        ctx = context_for(ai)
        ctx.rescale_rows(ai, ctx.prime0_for(ai.size(0), sp_prime_len), si, ri, round_at, exact)


def rescale_exact_rounding_fused(a, scales, rescaler, round_at, sp_prime_len):
    """In place on the kept rows (he_fused_cuda.cu:99-184)."""
    _rescale(a, scales, rescaler, round_at, sp_prime_len, True)


def rescale_non_exact_rounding_fused(a, scales, rescaler, sp_prime_len):
    _rescale(a, scales, rescaler, 0, sp_prime_len, False)


def switch_key_switch_later_part_extend(rns_len, state, l_enter, l_enter_start_offset, sp_prime_len):
    """ModUp extend of one digit group (he_fused_cuda.cu:276-355); single device, returns [rns_len, N]."""
    ctx = context_for(state)
    out = torch.empty(rns_len, state.size(1), dtype=state.dtype, device=state.device)
    le = l_enter if l_enter.numel() > 0 else None
    ctx.extend(rns_len, ctx.prime0_for(rns_len, sp_prime_len), state, le, l_enter_start_offset, out)
    return out


def codec_rotate_make_unsigned_reduce_2q(a, perm, _2q):
    outs = []
    for ai, pi, tq in zip(a, perm, _2q):
        ctx = context_for(ai)
        out = torch.empty_like(ai, memory_format=torch.contiguous_format)
        ctx.codec_rotate(ai, pi, tq, out)
        outs.append(out)
    return outs


def create_switcher_divide_by_p(c, p, PiRi):
    """ModDown (he_fused_cuda.cu:433-584).  The P_k^-1 tables come from the context (identical to the
    engine's PiRs, ckks_engine.py:201-239); `p` is modified in place by the chain-backward step exactly
    as in the reference."""
    outs = []
    for ci, pi in zip(c, p):
        ctx = context_for(ci)
        out = torch.empty_like(ci, memory_format=torch.contiguous_format)
        ctx.divide_by_p(ctx.num_ordinary - ci.size(0), ci, pi, out)
        outs.append(out)
    return outs


def mont_mult_sum_many_3d(*args, **kwargs):
    raise NotImplementedError("dead code in the reference (no Python caller); not provided")
