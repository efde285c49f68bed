"""tiberate/libs/wrapper/mont_ops.py mirror (schemas: csrc/ops/mont.cpp:138-155, mont_extra.cpp:67-79)."""

from __future__ import annotations

import torch

from .._native import ExplicitConsts
from . import context_for

MULT, ADD, SUB, ADD_R2Q, SUB_R2Q, ENTER_SCALAR, ENTER_RS, ENTER_RS_SCALE = range(8)
REDUCE, REDUCE_2Q, MAKE_SIGNED, MAKE_UNSIGNED, TILE_UNSIGNED, PC_ADD, ENTER_SCALAR_R2Q = range(8, 15)


def _binary(op, a, b, sp_prime_len):
    outs = []
    for ai, bi in zip(a, b):
        ctx = context_for(ai)
        out = torch.empty_like(ai, memory_format=torch.contiguous_format)
        ctx.pointwise(op, ai, bi, out, ctx.prime0_for(ai.size(0), sp_prime_len))
        outs.append(out)
    return outs


def _unary_inplace(op, a, sp_prime_len, scal=None):
    for i, ai in enumerate(a):
        ctx = context_for(ai)
        ctx.pointwise(op, ai, None, None, ctx.prime0_for(ai.size(0), sp_prime_len),
                      scal=None if scal is None else scal[i])


def mont_mult(a, b, sp_prime_len):
    return _binary(MULT, a, b, sp_prime_len)


def mont_add(a, b, sp_prime_len):
    return _binary(ADD, a, b, sp_prime_len)


def mont_sub(a, b, sp_prime_len):
    return _binary(SUB, a, b, sp_prime_len)


def mont_add_reduce_2q(a, b, sp_prime_len):
    return _binary(ADD_R2Q, a, b, sp_prime_len)


def mont_sub_reduce_2q(a, b, sp_prime_len):
    return _binary(SUB_R2Q, a, b, sp_prime_len)


def mont_enter_scalar(a, b, sp_prime_len):
    _unary_inplace(ENTER_SCALAR, a, sp_prime_len, scal=b)


def mont_enter_scalar_reduce_2q(a, b, sp_prime_len):
    """The reference's host wrapper returns its *input* tensors, not the computed ones
    (csrc/ops/cuda/mont_extra_cuda.cu:367), so callers observe a no-op; mirrored as is."""
    return a


def mont_enter_Rs(a, sp_prime_len):
    _unary_inplace(ENTER_RS, a, sp_prime_len)


def mont_enter_Rs_scale(a, sp_prime_len):
    _unary_inplace(ENTER_RS_SCALE, a, sp_prime_len)


def mont_reduce(a, sp_prime_len):
    _unary_inplace(REDUCE, a, sp_prime_len)


def reduce_2q(a, sp_prime_len):
    _unary_inplace(REDUCE_2Q, a, sp_prime_len)


def make_signed(a, sp_prime_len):
    _unary_inplace(MAKE_SIGNED, a, sp_prime_len)


def make_unsigned(a, sp_prime_len):
    _unary_inplace(MAKE_UNSIGNED, a, sp_prime_len)


def mont_enter(a, Rs, ql, qh, kl, kh):
    """Legacy form with explicit constant tensors (mont_cuda.cu:272-339): a <- MM(a, Rs_i), in place."""
    for ai, rs, l0, h0, l1, h1 in zip(a, Rs, ql, qh, kl, kh):
        ctx = context_for(ai)
        ec = ExplicitConsts(l0.data_ptr(), h0.data_ptr(), l1.data_ptr(), h1.data_ptr(), 0)
        ctx.pointwise(ENTER_SCALAR, ai, None, None, 0, scal=rs, ec=ec)


def mont_add_legacy(a, b, _2q):
    outs = []
    for ai, bi, tq in zip(a, b, _2q):
        ctx = context_for(ai)
        out = torch.empty_like(ai, memory_format=torch.contiguous_format)
        ctx.pointwise(ADD, ai, bi, out, 0, ec=ExplicitConsts(0, 0, 0, 0, tq.data_ptr()))
        outs.append(out)
    return outs


def tile_unsigned(a, _2q):
    """out[i][j] = a[j] + q_i (mont_cuda.cu:744-760)."""
    outs = []
    for ai, tq in zip(a, _2q):
        ctx = context_for(ai)
        out = torch.empty(tq.numel(), ai.size(-1), dtype=ai.dtype, device=ai.device)
        ctx.pointwise(TILE_UNSIGNED, ai.reshape(-1), None, out, 0, ec=ExplicitConsts(0, 0, 0, 0, tq.data_ptr()))
        outs.append(out)
    return outs


def _many(input, sp_prime_len, pairwise):
    outs = []
    for st in input:
        if st.dim() != 3:
            raise RuntimeError("Input must be 3D (K, C, N)")  # mont_extra_cuda.cu:144
        ctx = context_for(st)
        st = st.contiguous()
        out = torch.empty(st.size(1), st.size(2), dtype=st.dtype, device=st.device)
        ctx.add_many(st, out, ctx.prime0_for(st.size(1), sp_prime_len), pairwise)
        outs.append(out)
    return outs


def mont_add_many_3d(input, sp_prime_len):
    return _many(input, sp_prime_len, True)


def mont_reduce_add_many_3d(input, sp_prime_len):
    return _many(input, sp_prime_len, False)
